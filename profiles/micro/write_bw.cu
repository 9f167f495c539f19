// Write-bandwidth probe for the history materialisation: 60 rows x 10 M bytes.
#include <cstdio>
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
        cudaEventRecord(a); fill_rows_outer<<<148 * 8, 256>>>(p, row_vecs, T); cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
        printf("rows outer        %.1f us  %.0f GB/s\n", ms * 1e3, nvec * 16 / ms / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
