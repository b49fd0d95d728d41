// cub_anchor.cu -- TEST-SIDE comparator: cub::DeviceRadixSort::SortPairs
// (CCCL shipped with CUDA 12.9) against libdeltaq_cuda's own onesweep on the same (uint64 key, uint32 value) pairs.
// Not part of the product: nothing in deltaq_b200/ includes CUB.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/cub_anchor scripts/cub_anchor.cu -ldl
//   gpurun_out/cub_anchor deltaq_b200/libdeltaq_cuda.so
//
// Prints one JSON line per (count, key bits): milliseconds (best of 5, CUDA events) and "pass GB/s" = 24 B x count x
// ceil(bits/8) / time for both sorters -- the same algorithmic bytes bench.py's roofline uses -- and checks that both
// produce the same sorted keys and values (both sorts are stable).
#include <cub/cub.cuh>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                   \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

struct dq_ctx;
typedef int (*create_fn)(dq_ctx **, const int *, int);
typedef int (*sort_fn)(dq_ctx *, uint64_t *, uint32_t *, int32_t, int32_t, int32_t, int64_t *);

__global__ void fill(uint64_t *k, uint32_t *v, uint32_t n, int bits, uint64_t seed)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t x = (i + 1) * 0x9E3779B97F4A7C15ull + seed;  // splitmix64
        x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
        x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
        x ^= x >> 31;
        k[i] = bits == 64 ? x : (x & ((1ull << bits) - 1));
        v[i] = (uint32_t)i;
    }
}

__global__ void differ(const uint64_t *a, const uint64_t *b, const uint32_t *va, const uint32_t *vb, uint32_t n,
                       unsigned long long *bad)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        if (a[i] != b[i] || va[i] != vb[i]) atomicAdd(bad, 1ull);
}

int main(int argc, char **argv)
{
    const char *libpath = argc > 1 ? argv[1] : "deltaq_b200/libdeltaq_cuda.so";
    void *h = dlopen(libpath, RTLD_NOW);
    if (!h) {
        fprintf(stderr, "dlopen %s: %s\n", libpath, dlerror());
        return 1;
    }
    create_fn create = (create_fn)dlsym(h, "dq_cuda_create");
    sort_fn dq_sort = (sort_fn)dlsym(h, "dq_cuda_radix_sort_pairs_device");
    dq_ctx *ctx = nullptr;
    if (create(&ctx, nullptr, 0) != 0) return 1;

    const uint32_t counts[] = {1u << 20, 16777216u, 67108864u, 268435456u};
    const int bitsv[] = {64, 50};
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (uint32_t n : counts)
        for (int bits : bitsv) {
            uint64_t *k0, *ka, *kb, *kalt;
            uint32_t *v0, *va, *vb, *valt;
            unsigned long long *bad;
            CK(cudaMalloc(&k0, (size_t)n * 8)); CK(cudaMalloc(&ka, (size_t)n * 8)); CK(cudaMalloc(&kb, (size_t)n * 8)); CK(cudaMalloc(&kalt, (size_t)n * 8));
            CK(cudaMalloc(&v0, (size_t)n * 4)); CK(cudaMalloc(&va, (size_t)n * 4)); CK(cudaMalloc(&vb, (size_t)n * 4)); CK(cudaMalloc(&valt, (size_t)n * 4));
            CK(cudaMalloc(&bad, 8));
            fill<<<1024, 256>>>(k0, v0, n, bits, 7);
            size_t tmp_bytes = 0;
            cub::DoubleBuffer<uint64_t> dk(ka, kalt);
            cub::DoubleBuffer<uint32_t> dv(va, valt);
            CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, (int)n, 0, bits));
            void *tmp;
            CK(cudaMalloc(&tmp, tmp_bytes));
            float best_cub = 1e30f, best_dq = 1e30f;
            const uint64_t *cub_keys = nullptr;
            const uint32_t *cub_vals = nullptr;
            for (int it = 0; it < 6; ++it) {
                CK(cudaMemcpy(ka, k0, (size_t)n * 8, cudaMemcpyDeviceToDevice));
                CK(cudaMemcpy(va, v0, (size_t)n * 4, cudaMemcpyDeviceToDevice));
                cub::DoubleBuffer<uint64_t> k2(ka, kalt);
                cub::DoubleBuffer<uint32_t> v2(va, valt);
                CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(e0));
                CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, k2, v2, (int)n, 0, bits));
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                if (it) best_cub = ms < best_cub ? ms : best_cub;
                cub_keys = k2.Current();
                cub_vals = v2.Current();
            }
            for (int it = 0; it < 6; ++it) {
                CK(cudaMemcpy(kb, k0, (size_t)n * 8, cudaMemcpyDeviceToDevice));
                CK(cudaMemcpy(vb, v0, (size_t)n * 4, cudaMemcpyDeviceToDevice));
                CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(e0));
                if (dq_sort(ctx, kb, vb, (int32_t)n, 0, bits, nullptr) != 0) return 2;  // histogram pass + passes + sync
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                if (it) best_dq = ms < best_dq ? ms : best_dq;
            }
            CK(cudaMemset(bad, 0, 8));
            differ<<<1024, 256>>>(cub_keys, kb, cub_vals, vb, n, bad);
            unsigned long long hbad = 0;
            CK(cudaMemcpy(&hbad, bad, 8, cudaMemcpyDeviceToHost));
            const int passes = (bits + 7) / 8;
            const double bytes = 24.0 * n * passes;
            printf("{\"pairs\": %u, \"key_bits\": %d, \"passes\": %d, \"cub_ms\": %.4f, \"cub_pass_GBps\": %.1f, "
                   "\"deltaq_ms\": %.4f, \"deltaq_pass_GBps\": %.1f, \"deltaq_over_cub\": %.3f, \"mismatches\": %llu}\n",
                   n, bits, passes, best_cub, bytes / best_cub / 1e6, best_dq, bytes / best_dq / 1e6, best_cub / best_dq, hbad);
            fflush(stdout);
            cudaFree(k0); cudaFree(ka); cudaFree(kb); cudaFree(kalt); cudaFree(v0); cudaFree(va); cudaFree(vb); cudaFree(valt);
            cudaFree(bad); cudaFree(tmp);
        }
    return 0;
}
