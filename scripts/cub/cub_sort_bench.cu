// Comparison point only (not part of the product): cub::DeviceRadixSort on the same key shapes.
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdlib>
#include <vector>
int main(int argc, char **argv)
{
    size_t n = argc > 1 ? atoll(argv[1]) : 100000000;
    int begin_bit = argc > 2 ? atoi(argv[2]) : 0, end_bit = argc > 3 ? atoi(argv[3]) : 62;
    std::vector<unsigned long long> h(n);
    unsigned long long x = 88172645463325252ull;
    for (size_t i = 0; i < n; i++) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; h[i] = x >> 2; }
    unsigned long long *a, *b;
    cudaMalloc(&a, n * 8); cudaMalloc(&b, n * 8);
    void *tmp = nullptr; size_t tb = 0;
    cub::DoubleBuffer<unsigned long long> db(a, b);
    cub::DeviceRadixSort::SortKeys(tmp, tb, db, (int)n, begin_bit, end_bit);
    cudaMalloc(&tmp, tb);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int it = 0; it < 5; it++) {
        cudaMemcpy(a, h.data(), n * 8, cudaMemcpyHostToDevice);
        cub::DoubleBuffer<unsigned long long> d2(a, b);
        cudaEventRecord(e0);
        cub::DeviceRadixSort::SortKeys(tmp, tb, d2, (int)n, begin_bit, end_bit);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    int passes = (end_bit - begin_bit + 7) / 8;
    printf("cub SortKeys n=%zu bits[%d,%d): %.3f ms = %.2f Gkeys/s; ~%d passes -> %.3f ms/pass, %.0f GB/s per pass (2*8 B/key)\n", n, begin_bit, end_bit,
           best, n / best / 1e6, passes, best / passes, 16.0 * n / (best / passes) / 1e6);
    return 0;
}
