# builds A/B variants of the library that differ in the frustum kernels' compile-time shape (tools/probes/build/)
set -e
cd "$(dirname "$0")/../../wgpu-voxel-engine_b200/csrc"
make -s -j8
mkdir -p ../../tools/probes/build
build_variant() {   # name, defines
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC $2 -c kernels_frustum.cu -o ../../tools/probes/build/fr_$1.o
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/probes/build/libvox_$1.so build/api.o build/kernels_fill.o build/kernels_mesh.o ../../tools/probes/build/fr_$1.o build/kernels_peaks.o -cudart static
}
build_variant pair0 "-DVX_FR_PAIR_MIN=0"
build_variant pair1 "-DVX_FR_PAIR_MIN=1"
