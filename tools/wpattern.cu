// Micro-benchmark: write-only HBM bandwidth as a function of the store pattern (B200).
//   A  grid-stride: the whole grid sweeps the buffer front to back (memset-like)
//   B  private streams: each of G persistent CTAs writes its own contiguous region front to back
//      (the pattern of k_env: one tile of envs per CTA), 128 threads, 16 B per thread per store
//   C  like B, but CTAs take 140 KB tiles from a global counter (k_env_pipe's pattern)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/wpattern tools/wpattern.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void kA(int4* out, size_t n16) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
        __stcs(out + i, make_int4(1, 0, 1, 0));
}
__global__ void kB(int4* out, size_t n16) {
    const size_t per = n16 / gridDim.x;
    int4* base = out + per * blockIdx.x;
    for (size_t i = threadIdx.x; i < per; i += blockDim.x) __stcs(base + i, make_int4(1, 0, 1, 0));
}
__global__ void kC(int4* out, size_t n16, size_t tile16, unsigned* counter) {
    __shared__ unsigned t;
    const size_t ntiles = n16 / tile16;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) t = atomicAdd(counter, 1u);
        __syncthreads();
        if (t >= ntiles) break;
        int4* base = out + (size_t)t * tile16;
        for (size_t i = threadIdx.x; i < tile16; i += blockDim.x) __stcs(base + i, make_int4(1, 0, 1, 0));
    }
}

int main() {
    const size_t bytes = 131072ull * 8788ull / 16 * 16;
    const size_t n16 = bytes / 16;
    int4* buf; unsigned* ctr;
    cudaMalloc(&buf, bytes); cudaMalloc(&ctr, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char* name, auto launch) {
        float best = 1e9f;
        for (int r = 0; r < 12; r++) {
            cudaMemset(ctr, 0, 4);
            cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (r >= 2 && ms < best) best = ms;
        }
        printf("%-40s %.4f ms  %.0f GB/s\n", name, best, bytes / best / 1e6);
    };
    for (int thr : {128, 256, 1024}) {
        char nm[64];
        snprintf(nm, 64, "A grid-stride %d thr x %d CTAs", thr, 148 * 2048 / thr);
        run(nm, [&] { kA<<<148 * 2048 / thr, thr>>>(buf, n16); });
    }
    for (int g : {148, 592, 1184, 2368}) {
        char nm[64];
        snprintf(nm, 64, "B private streams, %d CTAs x 128 thr", g);
        run(nm, [&] { kB<<<g, 128>>>(buf, n16); });
    }
    for (size_t tileKB : {35, 70, 140, 281}) {
        char nm[64];
        snprintf(nm, 64, "C counter tiles of %zu KB, 1184x128", tileKB);
        run(nm, [&] { kC<<<1184, 128>>>(buf, n16, tileKB * 1024 / 16, ctr); });
    }
    for (int g : {148, 296}) {
        char nm[64];
        snprintf(nm, 64, "C tiles 140 KB, %d CTAs x 1024 thr", g);
        run(nm, [&] { kC<<<g, 1024>>>(buf, n16, 140 * 1024 / 16, ctr); });
    }
    cudaError_t err = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(err));
    return err != cudaSuccess;
}
