// Times the REAL k32_cols / k32_rows kernels (included from the library source) in the microbench harness.
#include "../../style_transfer_based_holographic_imaging_b200/csrc/asm_b200.cu"
#include <cstdio>
int main() {
    using namespace asmb;
    const int R = 6, L = 1024, nimg_in = 96;
    float2 *ws, *in, *tw; double* kzt; float* z; float* out;
    cudaMalloc(&ws, (size_t)R * L * L * 8); cudaMalloc(&in, (size_t)nimg_in * L * L * 8); cudaMalloc(&out, (size_t)nimg_in * L * L * 4);
    cudaMalloc(&tw, 4096 * 8); cudaMalloc(&kzt, 513 * L * 8); cudaMalloc(&z, 4096 * 4);
    cudaMemset(ws, 0, (size_t)R * L * L * 8); cudaMemset(in, 0, (size_t)nimg_in * L * L * 8); cudaMemset(z, 0, 4096 * 4);
    Params p{};
    p.in0 = in; p.out0 = out; p.z = z; p.ws = ws; p.tw = tw; p.kzt = kzt; p.s2 = 1e-7; p.lambda = 532e-9; p.inv_lambda = 1 / 532e-9;
    p.inv_m2 = 1.f / (1024.f * 1024.f); p.planes = nimg_in; p.C = 1; p.N = L; p.M = L; p.P = 0; p.in_mode = 0; p.out_mode = 1;
    k32_setup<<<296, 256>>>(tw, kzt, nullptr, 0, p.s2, p.inv_lambda * 0.159);
    const size_t smem_rows = (size_t)K32_ROW_WARPS * K32_LP * 8 + (size_t)K32_TW * 8;
    const size_t smem_cols = (size_t)K32_SLAB_ROWS * K32_CC * 8 + (size_t)(L / 2 + 1) * K32_CC * 8 + (size_t)K32_TW * 8 + 2 * K32_CC * 8;
    cudaFuncSetAttribute(k32_rows_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows);
    cudaFuncSetAttribute(k32_rows_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows);
    cudaFuncSetAttribute(k32_cols, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cols);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto timeit = [&](auto f, const char* name, int imgs) {
        f(); cudaDeviceSynchronize();
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("%-40s %.3f ms = %.2f us/image  %s\n", name, ms, ms * 1e3 / imgs, cudaGetErrorString(cudaGetLastError()));
    };
    const int reps = 16;
    timeit([&] { for (int r = 0; r < reps; ++r) k32_cols<<<296, 256, smem_cols>>>(p, 0, R); }, "real k32_cols, 6 img x16 launches", R * reps);
    timeit([&] { for (int r = 0; r < reps; ++r) k32_rows_inv<<<296, 256, smem_rows>>>(p, 0, R * L); }, "real k32_rows_inv (|U|^2), 6 img x16", R * reps);
    // rows_fwd writes ws[img] for img < nimg: use 6 images of input per launch
    timeit([&] { for (int r = 0; r < reps; ++r) k32_rows_fwd<<<296, 256, smem_rows>>>(p, (r * R) % nimg_in, R * L); }, "real k32_rows_fwd, 6 img x16", R * reps);
    return 0;
}
