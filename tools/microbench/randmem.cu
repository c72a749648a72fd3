// randmem.cu -- what can HBM3e / L2 of one B200 sustain for the access shapes of k_step?  (tuning aid, not product code)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a randmem.cu -o randmem && ./randmem
// Shapes: streaming copy; one random 32-byte sector per thread (read / write); a random 4 KB-strided chunk of 17 x 32 B
// read 8 B per sector (the selection scan); a returning 64-bit atomic on a random word of an 8 MB (L2-resident) or
// 4 GB region (the inbox slot reservation).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t z) { z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
__global__ void k_stream(const uint4 *a, uint4 *b, size_t n) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i]; }
__global__ void k_rand_read(const uint4 *a, size_t nsec, uint64_t seed, unsigned long long *sink)
{
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t s = mix(seed + t) % nsec;
    const uint4 v = a[2 * s]; const uint4 w = a[2 * s + 1];
    if ((v.x ^ w.y) == 0x12345678u) atomicAdd(sink, 1ull);
}
__global__ void k_rand_write(uint4 *a, size_t nsec, uint64_t seed)
{
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t s = mix(seed + t) % nsec;
    const uint4 v = make_uint4((unsigned)t, 1, 2, 3);
    a[2 * s] = v; a[2 * s + 1] = v;
}
__global__ void k_chunk_scan(const double *a, size_t nchunk, uint64_t seed, unsigned long long *sink, int per)
{
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t c = mix(seed + t) % nchunk;          // 4 KB chunk
    const double *p = a + c * 512 + (mix(seed ^ t) & 63) * 4;    // start slot inside the ring
    double m = 1e300;
    for (int k = 0; k < per; ++k) { const double v = p[(k * 4) & 511]; m = v < m ? v : m; }
    if (m == 123.0) atomicAdd(sink, 1ull);
}
__global__ void k_chunk_seq(const double *a, size_t nchunk, unsigned long long *sink, int per)     // LP i -> chunk i (what k_step does)
{
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= nchunk) return;
    const double *p = a + t * 512 + (mix(t) & 63) * 4;
    double m = 1e300;
    for (int k = 0; k < per; ++k) { const double v = p[(k * 4) & 511]; m = v < m ? v : m; }
    if (m == 123.0) atomicAdd(sink, 1ull);
}
__global__ void k_rand_atomic(unsigned long long *a, size_t nword, uint64_t seed, unsigned long long *sink)
{
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t s = mix(seed + t) % nword;
    const unsigned long long old = atomicAdd(&a[s], 1ull << 32);
    if (old == 0x123456789ull) atomicAdd(sink, 1ull);
}
template <class F> float timeit(F f, int reps = 5)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); for (int i = 0; i < reps; ++i) f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / reps;
}
int main()
{
    const size_t bytes = 4ull << 30; const size_t nsec = bytes / 32;
    uint4 *a, *b; unsigned long long *sink, *small;
    cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&sink, 8); cudaMalloc(&small, 8 << 20);
    cudaMemset(a, 1, bytes); cudaMemset(b, 2, bytes); cudaMemset(small, 0, 8 << 20);
    const int T = 256; uint64_t seed = 1;
    float ms = timeit([&] { k_stream<<<148 * 16, T>>>(a, b, bytes / 16); });
    printf("stream copy 4 GB            : %8.1f us  %7.1f GB/s (read+write)\n", ms * 1e3, 2.0 * bytes / ms / 1e6);
    for (size_t n : {1ull << 20, 1ull << 22, 1ull << 24}) {
        ms = timeit([&] { k_rand_read<<<n / T, T>>>(a, nsec, seed++, sink); });
        printf("random 32 B reads  n=%8zu: %8.1f us  %7.1f GB/s  %6.2f G sectors/s\n", n, ms * 1e3, 32.0 * n / ms / 1e6, n / ms / 1e6);
        ms = timeit([&] { k_rand_write<<<n / T, T>>>(a, nsec, seed++); });
        printf("random 32 B writes n=%8zu: %8.1f us  %7.1f GB/s  %6.2f G sectors/s\n", n, ms * 1e3, 32.0 * n / ms / 1e6, n / ms / 1e6);
    }
    for (int per : {8, 17, 32}) {
        const size_t n = 1 << 20;
        ms = timeit([&] { k_chunk_scan<<<n / T, T>>>((const double *)a, bytes / 4096, seed++, sink, per); });
        printf("random chunks, %2d sectors/thread, n=1M: %8.1f us  %7.1f GB/s (sector bytes)\n", per, ms * 1e3, 32.0 * per * n / ms / 1e6);
        ms = timeit([&] { k_chunk_seq<<<n / T, T>>>((const double *)a, bytes / 4096, sink, per); });
        printf("LP-ordered chunks, %2d sectors/thread, n=1M: %8.1f us  %7.1f GB/s (sector bytes)\n", per, ms * 1e3, 32.0 * per * n / ms / 1e6);
    }
    for (size_t n : {1ull << 20, 1ull << 22}) {
        ms = timeit([&] { k_rand_atomic<<<n / T, T>>>(small, (8u << 20) / 8, seed++, sink); });
        printf("returning atomics, 8 MB region, n=%8zu: %8.1f us  %6.2f G atomics/s\n", n, ms * 1e3, n / ms / 1e6);
        ms = timeit([&] { k_rand_atomic<<<n / T, T>>>((unsigned long long *)a, bytes / 8, seed++, sink); });
        printf("returning atomics, 4 GB region, n=%8zu: %8.1f us  %6.2f G atomics/s\n", n, ms * 1e3, n / ms / 1e6);
    }
    return 0;
}
