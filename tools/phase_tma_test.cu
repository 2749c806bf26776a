// Stand-alone check of the 5-D "phase view" TMA stores used for OUT_PHASE outputs (conv_igemm_kernel, OutDesc::tma == 2).
// Findings on B200: a box that overruns the tensor extent is clipped; a NEGATIVE start coordinate raises "illegal
// instruction"; the 64-byte swizzle of the shared-memory source is a function of the ADDRESS (a source that starts k x 128
// bytes into a 1024-byte aligned staged chunk is read correctly).
// nvcc -gencode arch=compute_100a,code=sm_100a -o tools/phase_tma_test tools/phase_tma_test.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstdint>
__global__ void k(const __grid_constant__ CUtensorMap map, int row_off, int c0, int xh, int yh, int j) {
    __shared__ __align__(1024) uint16_t buf[128 * 32];
    for (int i = threadIdx.x; i < 128 * 32; i += blockDim.x) {
        const int row = i / 32, ch = i % 32;
        const int slot = (ch / 8) ^ ((row >> 1) & 3);         // 64-byte swizzle by absolute row
        buf[row * 32 + slot * 8 + (ch % 8)] = (uint16_t)(row * 32 + ch + 1);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t s = (uint32_t)__cvta_generic_to_shared(buf) + (uint32_t)row_off * 64u;
        asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(reinterpret_cast<uint64_t>(&map)),
                     "r"(s), "r"(c0), "r"(0), "r"(xh), "r"(yh), "r"(j) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}
int main() {
    const int nmax = 2, H = 40, W = 40, C = 64;      // producer output 40x40 -> padded 42x42, 21 pairs per row; planes 22 x 22
    const uint64_t pw = W / 2 + 2, ph = H / 2 + 2, plane = ph * pw, eb = C * 2;
    const size_t elems = 4 * nmax * plane * C;
    uint16_t* d; cudaMalloc(&d, elems * 2);
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                           CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fp; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    int total_bad = 0;
    const int cases[][6] = {   // box pairs, staged row offset, c0, xh, yh, j
        {64, 0, 32, 3, 0, 4}, {64, 36, 0, 0, 1, 4}, {16, 0, 0, 0, 2, 1}, {16, 10, 32, 5, 3, 0}, {4, 6, 0, 17, 4, 5}, {1, 126, 32, 20, 5, 0}, {2, 124, 0, 0, 6, 1}, {8, 2, 0, 19, 7, 1}};
    for (auto& cs : cases) {
        const int boxx = cs[0], roff = cs[1], c0 = cs[2], xh0 = cs[3], yh0 = cs[4], j0 = cs[5];
        CUtensorMap m;
        cuuint64_t dims[5] = {(cuuint64_t)C, 2, (cuuint64_t)((W + 2) / 2), (cuuint64_t)((H + 2) / 2), (cuuint64_t)(3 * nmax)};
        cuuint64_t strides[4] = {nmax * plane * eb, eb, pw * eb, plane * eb};
        cuuint32_t box[5] = {32, 2, (cuuint32_t)boxx, 1, 1}, es[5] = {1, 1, 1, 1, 1};
        CUresult r = ((Fn)fp)(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r) { printf("encode box %d failed %d\n", boxx, (int)r); return 1; }
        cudaMemset(d, 0, elems * 2);
        k<<<1, 128>>>(m, roff, c0, xh0, yh0, j0);
        cudaError_t e = cudaDeviceSynchronize();
        if (e) { printf("box %d roff %d xh %d: %s\n", boxx, roff, xh0, cudaGetErrorString(e)); return 2; }
        std::vector<uint16_t> hb(elems);
        cudaMemcpy(hb.data(), d, elems * 2, cudaMemcpyDeviceToHost);
        std::vector<uint16_t> ex(elems, 0);
        for (int x = 0; x < boxx; ++x)
            for (int cp = 0; cp < 2; ++cp) {
                const int xh = xh0 + x;
                if (xh >= (W + 2) / 2) continue;
                const int row = roff + 2 * x + cp;
                const size_t pix = ((size_t)cp * nmax + j0) * plane + (size_t)yh0 * pw + xh;
                for (int ch = 0; ch < 32; ++ch) ex[pix * C + c0 + ch] = (uint16_t)(row * 32 + ch + 1);
            }
        int bad = 0;
        for (size_t i = 0; i < elems; ++i) if (ex[i] != hb[i]) { if (bad < 4) printf("  mismatch @%zu got %d expect %d\n", i, hb[i], ex[i]); ++bad; }
        printf("box %2d pairs, staged row offset %3d, xh %2d: mismatches %d\n", boxx, roff, xh0, bad);
        total_bad += bad;
    }
    printf(total_bad ? "FAILED\n" : "all phase-view stores as expected\n");
    return total_bad != 0;
}
