// SASS instruction-count probe for the K0 field primitives:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -cubin -o /tmp/p.cubin tools/sass_probe.cu && cuobjdump -sass /tmp/p.cubin
#include "../snark_verifier_axiom_b200/csrc/field.cuh"
__global__ void k_probe_mul(Fq* a) { a[0] = a[1] * a[2]; }
__global__ void k_probe_sqr(Fq* a) { a[0] = a[1].sqr(); }
__global__ void k_probe_dot2(Fq* a) { a[0] = Fq::dot2(a[1], a[2], a[3], a[4]); }
__global__ void k_probe_inv(Fq* a) { a[threadIdx.x] = a[threadIdx.x + 64].inv(); }
