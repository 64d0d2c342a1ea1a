// Micro-benchmark: latency of __match_any_sync as a function of the number of distinct values in the warp, against the
// shared-memory alternative (read cursor, atomicAdd, read again).  nvcc -arch=sm_100a -o match_bench match_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_match(int distinct, int iters, long long *out, unsigned *sink) {
    int v = (threadIdx.x & 31) % distinct;
    unsigned acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        unsigned g = __match_any_sync(0xffffffffu, v);
        acc += g;
        v = (v + (g & 1)) % distinct + 0 * i;  // dependent chain
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0) / iters;
    sink[threadIdx.x] = acc;
}
__global__ void k_smem(int distinct, int iters, long long *out, unsigned *sink) {
    __shared__ int cur[2048];
    for (int i = threadIdx.x; i < 2048; i += 32) cur[i] = 0;
    __syncwarp();
    int v = ((threadIdx.x & 31) * 37) % distinct;
    unsigned acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        const int base = cur[v];
        __syncwarp();
        const int pos = atomicAdd(&cur[v], 1);
        __syncwarp();
        const int k = cur[v] - base;
        const unsigned coll = __ballot_sync(0xffffffffu, k > 1);
        int rank = 0;
        if (k > 1) rank = __popc(__match_any_sync(__activemask(), v) & ((1u << (threadIdx.x & 31)) - 1u));
        __syncwarp();
        acc += base + rank + pos + coll;
        v = (v + (acc & 1)) % distinct;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0) / iters;
    sink[threadIdx.x] = acc;
}
int main() {
    long long *d_out, h;
    unsigned *sink;
    cudaMalloc(&d_out, 8);
    cudaMalloc(&sink, 4 * 32);
    int ds[] = {1, 2, 4, 8, 16, 24, 32};
    for (int d : ds) {
        k_match<<<1, 32>>>(d, 2000, d_out, sink);
        cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
        long long m = h;
        k_smem<<<1, 32>>>(d == 32 ? 2048 : d, 2000, d_out, sink);
        cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
        printf("distinct %2d: match.any %lld cycles/iter (incl. ~30 of chain), smem read+atomic+read %lld cycles/iter\n", d, m, h);
    }
    return 0;
}
