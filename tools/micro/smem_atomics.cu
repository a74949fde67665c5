// Microbenchmark: shared-memory atomic throughput on sm_100a (spread addresses, hashed), per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_atomics smem_atomics.cu && ./smem_atomics
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
__device__ __forceinline__ u32 mix(u32 x) { x *= 0x9E3779B1u; x ^= x >> 15; x *= 0x85EBCA6Bu; x ^= x >> 13; return x; }
struct __align__(16) K2 { u64 lo, hi; };
__device__ __forceinline__ K2 cas128(K2* addr, K2 cmp, K2 val) {
    K2 old; u32 sa = (u32)__cvta_generic_to_shared(addr);
    asm volatile("{\n .reg .b128 c, s, d;\n mov.b128 c, {%3, %4};\n mov.b128 s, {%5, %6};\n atom.shared.cas.b128 d, [%2], c, s;\n mov.b128 {%0, %1}, d;\n}\n"
                 : "=l"(old.lo), "=l"(old.hi) : "r"(sa), "l"(cmp.lo), "l"(cmp.hi), "l"(val.lo), "l"(val.hi) : "memory");
    return old;
}
template <int MODE>
__global__ void __launch_bounds__(512) k(int iters, u64* out) {
    extern __shared__ __align__(16) unsigned char sm[];
    u32* s32 = (u32*)sm; u64* s64 = (u64*)sm; K2* s128 = (K2*)sm;
    const int N32 = 16384, N64 = 8192, N128 = 4096;
    for (int i = threadIdx.x; i < N32; i += blockDim.x) s32[i] = 0;
    __syncthreads();
    u32 h = mix(threadIdx.x + blockIdx.x * 977u);
    u64 acc = 0;
    for (int it = 0; it < iters; it++) {
        h = mix(h + it);
        if (MODE == 0) atomicAdd(&s32[h & (N32 - 1)], 1u);                       // RED-like add, no return
        if (MODE == 1) acc += atomicAdd(&s32[h & (N32 - 1)], 1u);                // add with return
        if (MODE == 2) atomicOr(&s32[h & (N32 - 1)], h);                         // or, no return
        if (MODE == 3) acc += atomicCAS(&s64[h & (N64 - 1)], 0ull, (u64)h);      // cas64
        if (MODE == 4) { K2 o = cas128(&s128[h & (N128 - 1)], K2{0, 0}, K2{h, h}); acc += o.lo; }  // cas128
        if (MODE == 5) acc += s32[h & (N32 - 1)];                                // plain LDS.32
        if (MODE == 6) acc += s64[h & (N64 - 1)];                                // plain LDS.64
        if (MODE == 7) s32[h & (N32 - 1)] = h;                                   // plain STS.32
        if (MODE == 8) acc += atomicAdd(&s64[h & (N64 - 1)], 1ull);              // add64 with return
        if (MODE == 9) atomicAdd(&s64[h & (N64 - 1)], 1ull);                     // add64 no return
    }
    if (acc == 0x1234567) out[0] = acc;
}
template <int MODE> void run(const char* name, int threads, int ctas_per_sm) {
    int sms = 148, iters = 4000;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    u64* d; cudaMalloc(&d, 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<sms * ctas_per_sm, threads, 65536>>>(100, d);
    cudaEventRecord(a);
    k<MODE><<<sms * ctas_per_sm, threads, 65536>>>(iters, d);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double ops_per_sm = (double)iters * threads * ctas_per_sm;
    double cyc = ms * 1e-3 * 1.965e9;
    printf("%-22s thr=%4d x%d: %8.3f ms  %6.2f cycles per lane-op per SM (%.1f Gop/s chip)\n", name, threads, ctas_per_sm, ms, cyc / ops_per_sm,
           ops_per_sm * sms / ms / 1e6);
    cudaFree(d);
}
int main() {
    for (int c = 1; c <= 2; c++) {
        run<0>("atomicAdd32 noret", 512, c); run<1>("atomicAdd32 ret", 512, c); run<2>("atomicOr32 noret", 512, c);
        run<3>("atomicCAS64", 512, c); run<4>("atomicCAS128", 512, c); run<8>("atomicAdd64 ret", 512, c); run<9>("atomicAdd64 noret", 512, c);
        run<5>("LDS.32", 512, c); run<6>("LDS.64", 512, c); run<7>("STS.32", 512, c);
    }
    return 0;
}
