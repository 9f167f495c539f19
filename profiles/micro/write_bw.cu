// Write-bandwidth probe for the history materialisation: 60 rows x 10 M bytes.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void fill_linear(uint4 *p, size_t nvec)
{
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < nvec; g += (size_t)gridDim.x * blockDim.x) __stcs(p + g, z);
}
__global__ void fill_linear_plain(uint4 *p, size_t nvec)
{
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < nvec; g += (size_t)gridDim.x * blockDim.x) p[g] = z;
}
__global__ void fill_columns(uint4 *p, size_t row_vecs, int T)
{
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < row_vecs; g += (size_t)gridDim.x * blockDim.x)
        for (int t = 0; t < T; ++t) __stcs(p + (size_t)t * row_vecs + g, z);
}
__global__ void fill_rows_outer(uint4 *p, size_t row_vecs, int T)
{
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int t = 0; t < T; ++t)
        for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < row_vecs; g += (size_t)gridDim.x * blockDim.x)
            __stcs(p + (size_t)t * row_vecs + g, z);
}
// TMA bulk stores (cp.async.bulk shared -> global): every warp owns `chunk` bytes of every row; one lane issues the row's copy
template <int CHUNK>
__global__ void fill_columns_bulk(uint8_t *p, size_t row_bytes, int T)
{
    extern __shared__ __align__(128) uint8_t tile[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    uint8_t *mine = tile + (size_t)wib * CHUNK;
    for (int k = lane * 16; k < CHUNK; k += 512) *reinterpret_cast<uint4 *>(mine + k) = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const size_t chunks = row_bytes / CHUNK;
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(mine);
    for (size_t g = (size_t)blockIdx.x * wpb + wib; g < chunks; g += (size_t)gridDim.x * wpb) {
        if (lane == 0) {
            for (int t = 0; t < T; ++t) {
                uint8_t *dst = p + (size_t)t * row_bytes + g * CHUNK;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(saddr), "r"(CHUNK) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// the same with the stores spread over the 32 lanes (lane = rows t, t+32, ...)
template <int CHUNK>
__global__ void fill_columns_bulk_lanes(uint8_t *p, size_t row_bytes, int T)
{
    extern __shared__ __align__(128) uint8_t tile[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    uint8_t *mine = tile + (size_t)wib * CHUNK;
    for (int k = lane * 16; k < CHUNK; k += 512) *reinterpret_cast<uint4 *>(mine + k) = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const size_t chunks = row_bytes / CHUNK;
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(mine);
    for (size_t g = (size_t)blockIdx.x * wpb + wib; g < chunks; g += (size_t)gridDim.x * wpb) {
        for (int t = lane; t < T; t += 32) {
            uint8_t *dst = p + (size_t)t * row_bytes + g * CHUNK;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(saddr), "r"(CHUNK) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// st.cs with a warp owning 512 B of every row (what seir_materialize_kernel does), rows inner
__global__ void fill_warp_columns(uint4 *p, size_t row_vecs, int T)
{
    const uint4 z = make_uint4(0, 0, 0, 0);
    const int lane = threadIdx.x & 31;
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5, warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t g = warp; g < row_vecs / 32; g += warps)
        for (int t = 0; t < T; ++t) __stcs(p + (size_t)t * row_vecs + g * 32 + lane, z);
}
int main()
{
    const int T = 60; const size_t n = 10002432; const size_t row_vecs = n / 16, nvec = row_vecs * T;
    uint4 *p; cudaMalloc(&p, nvec * 16);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float ms;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(a); cudaMemsetAsync(p, 0, nvec * 16); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        printf("cudaMemset        %.1f us  %.0f GB/s\n", ms * 1e3, nvec * 16 / ms / 1e6);
        for (int blocks : {148 * 4, 148 * 8, 148 * 16, 148 * 32}) {
            cudaEventRecord(a); fill_linear<<<blocks, 256>>>(p, nvec); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
            printf("linear stcs  %5d %.1f us  %.0f GB/s\n", blocks, ms * 1e3, nvec * 16 / ms / 1e6);
        }
        cudaEventRecord(a); fill_linear_plain<<<148 * 16, 256>>>(p, nvec); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        printf("linear plain      %.1f us  %.0f GB/s\n", ms * 1e3, nvec * 16 / ms / 1e6);
        cudaEventRecord(a); fill_columns<<<148 * 8, 256>>>(p, row_vecs, T); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        printf("columns (60 rows per thread) %.1f us  %.0f GB/s\n", ms * 1e3, nvec * 16 / ms / 1e6);
        for (int blocks : {148 * 4, 148 * 8, 148 * 16}) {
            cudaEventRecord(a); fill_warp_columns<<<blocks, 256>>>(p, row_vecs, T); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
            printf("warp columns stcs %5d %.1f us  %.0f GB/s\n", blocks, ms * 1e3, nvec * 16 / ms / 1e6);
            cudaEventRecord(a); fill_columns_bulk<512><<<blocks, 256, 8 * 512>>>((uint8_t *)p, n, T); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
            printf("bulk 512 B, one lane %5d %.1f us  %.0f GB/s\n", blocks, ms * 1e3, nvec * 16 / ms / 1e6);
            cudaEventRecord(a); fill_columns_bulk_lanes<512><<<blocks, 256, 8 * 512>>>((uint8_t *)p, n, T); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
            printf("bulk 512 B, lanes    %5d %.1f us  %.0f GB/s\n", blocks, ms * 1e3, nvec * 16 / ms / 1e6);
            cudaEventRecord(a); fill_columns_bulk_lanes<2048><<<blocks, 256, 8 * 2048>>>((uint8_t *)p, n, T); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
            printf("bulk 2048 B, lanes   %5d %.1f us  %.0f GB/s\n", blocks, ms * 1e3, nvec * 16 / ms / 1e6);
            cudaEventRecord(a); fill_columns_bulk_lanes<4096><<<blocks, 256, 8 * 4096>>>((uint8_t *)p, n, T); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
            printf("bulk 4096 B, lanes   %5d %.1f us  %.0f GB/s\n", blocks, ms * 1e3, nvec * 16 / ms / 1e6);
        }
        cudaEventRecord(a); fill_rows_outer<<<148 * 8, 256>>>(p, row_vecs, T); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        printf("rows outer        %.1f us  %.0f GB/s\n", ms * 1e3, nvec * 16 / ms / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
