"""cube.rs:16-21 CubeType and texture.rs:238-243 TEXTURE_INDEX_LIST."""
GRASS, DIRT, AIR = 0, 1, 2
# Rotation, in the per-cube emission order of chunk.rs:77-240
LEFT, RIGHT, BACK, FRONT, DOWN, UP = range(6)
TEXTURE_INDEX_LIST = ((0, 0, 0, 0, 2, 1), (1, 1, 1, 1, 1, 1))
