bash tools/prof.sh f3city --map city10k --kind long --n 16777216
bash tools/prof.sh f3rooms --map rooms8 --kind uniform --n 16777216
bash tools/prof.sh f3mega --map mega200k --kind short --n 16777216
