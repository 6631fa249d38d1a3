/* oracle shim: glm/gtc/type_ptr.hpp -- everything lives in glm/glm.hpp */
