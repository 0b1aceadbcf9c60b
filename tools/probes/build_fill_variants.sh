# builds A/B variants of the library that differ in compile-time knobs of the lattice fill kernel (tools/probes/build/)
set -e
cd "$(dirname "$0")/../../wgpu-voxel-engine_b200/csrc"
make -s -j8
mkdir -p ../../tools/probes/build
build_variant() {   # name, defines
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -fmad=false $2 -Xptxas -v -c kernels_fill.cu -o ../../tools/probes/build/fill_$1.o 2>&1 | grep -A1 "vx_fill_lattice_kernelILb0ELi4" | grep -E "registers|spill" | head -2
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/probes/build/libvoxf_$1.so build/api.o ../../tools/probes/build/fill_$1.o build/kernels_mesh.o build/kernels_frustum.o build/kernels_peaks.o -cudart static
}
build_variant it2 ""
build_variant pu "-DVX_FILL_PAIRED_U=1"
build_variant puc3 "-DVX_FILL_PAIRED_U=1 -DVX_FILL_MIN_CTAS=3"
