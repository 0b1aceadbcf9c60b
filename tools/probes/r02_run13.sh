cd tools/probes
./overlap_probe 30720 49152 1 0 8
./overlap_probe 30720 49152 1 1 8
./overlap_probe 30720 49152 0 1 8
./overlap_probe 30720 116736 1 1 2
./overlap_probe 30720 81920 1 1 4
