set -e
cd "$(dirname "$0")/../../wgpu-voxel-engine_b200/csrc"
make -s -j8
mkdir -p ../../tools/probes/build
build_variant() {
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC $2 -Xptxas -v -c kernels_mesh.cu -o ../../tools/probes/build/mesh_$1.o 2>&1 | grep -A1 "vx_mesh_emit_kernelILb0ELi1ELi0" | grep -E "registers|spill" | head -2
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/probes/build/libvoxm_$1.so build/api.o build/kernels_fill.o ../../tools/probes/build/mesh_$1.o build/kernels_frustum.o build/kernels_peaks.o -cudart static
}
build_variant u1 "-DVX_EMIT_TILE_UNROLL=1"
build_variant u2 "-DVX_EMIT_TILE_UNROLL=2"
build_variant u4 "-DVX_EMIT_TILE_UNROLL=4"
