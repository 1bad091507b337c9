// Does cp.async.bulk.prefetch.L2 really bring data into L2 on B200?  prefetch_kernel pulls `bytes` in chunks of `chunk`
// bytes (one thread per chunk), read_kernel then reads them; compare read_kernel's time / DRAM bytes with and without.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pf_test tools/pf_test.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
__global__ void prefetch_kernel(const char *p, size_t bytes, size_t chunk, int hint) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t off = i * chunk;
    if (off >= bytes) return;
    const uint32_t n = (uint32_t)(off + chunk <= bytes ? chunk : bytes - off);
    if (hint) asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(p + off), "r"(n), "l"(pol) : "memory");
    else asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p + off), "r"(n) : "memory");
}
__global__ void read_kernel(const uint4 *p, size_t n16, uint4 *sink) {
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = p[i]; acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
    }
    if (acc.x == 0x12345678u) *sink = acc;
}
__global__ void flush_kernel(uint4 *p, size_t n16) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) p[i] = make_uint4(i, 1, 2, 3);
}
int main(int argc, char **argv) {
    const size_t bytes = (size_t)(argc > 1 ? atoi(argv[1]) : 32) << 20;
    char *buf; uint4 *sink, *junk; const size_t junk_bytes = (size_t)1 << 30;
    cudaMalloc(&buf, bytes); cudaMalloc(&sink, 16); cudaMalloc(&junk, junk_bytes);
    cudaMemset(buf, 1, bytes);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (size_t chunk : {(size_t)0, (size_t)4096, (size_t)65536, (size_t)147456, (size_t)1048576}) {
        for (int hint = 0; hint < 2; ++hint) {
            float best = 1e9;
            for (int it = 0; it < 3; ++it) {
                flush_kernel<<<1184, 256>>>(junk, junk_bytes / 16);          // evict everything
                if (chunk) {
                    const size_t n = (bytes + chunk - 1) / chunk;
                    prefetch_kernel<<<(unsigned)((n + 63) / 64), 64>>>(buf, bytes, chunk, hint);
                }
                cudaDeviceSynchronize();
                // give the TMA prefetches time to land
                for (volatile int s = 0; s < 1000000; ++s) {}
                cudaEventRecord(a);
                read_kernel<<<1184, 256>>>((const uint4 *)buf, bytes / 16, sink);
                cudaEventRecord(b); cudaEventSynchronize(b);
                float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
            }
            printf("%zu MB, prefetch chunk %8zu B hint %d: read %.3f ms = %.0f GB/s (%s)\n", bytes >> 20, chunk, hint, best,
                   bytes / (best * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
            if (!chunk) break;
        }
    }
    return 0;
}
