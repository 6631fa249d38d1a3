/* oracle shim: glm/gtc/matrix_transform.hpp -- everything lives in glm/glm.hpp */
