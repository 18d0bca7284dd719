// Developer probe (not part of the product library): runs a host-described sequence of tcgen05.mma
// instructions on host-built shared-memory images and returns the raw TMEM accumulator, plus the
// clock64 time of the sequence.  Used to pin down, on the B200 itself, the operand layouts the
// tensor-core mel projection relies on (MN-major A without swizzle at a free stride, K-major B
// chunks, M = 64 accumulator placement, column-window accumulation) and to measure what a
// DFT-as-GEMM stage would cost on the tensor pipe.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -shared \
//        -o tools/libumma_probe.so tools/umma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct MmaOp {            // one tcgen05.mma
    uint32_t a_off16;     // added to the A descriptor's start-address field (16-byte units)
    uint32_t b_off16;     // same for B
    uint32_t d_col;       // TMEM column offset; bit 31 = accumulate (0: overwrite D)
    uint32_t idesc;       // instruction descriptor (upper 32 bits of the CUTLASS idescE)
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const uint4* __restrict__ a_img, int a_vec, const uint4* __restrict__ b_img, int b_vec,
             const MmaOp* __restrict__ ops, int n_ops, uint64_t a_desc_hi, uint64_t b_desc_hi,
             int reps, int tmem_cols, float* __restrict__ d_out, int out_cols, long long* __restrict__ cycles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint4* A = reinterpret_cast<uint4*>(smem);
    uint4* B = A + a_vec;
    MmaOp* T = reinterpret_cast<MmaOp*>(B + b_vec);
    uint64_t* bar = reinterpret_cast<uint64_t*>(T + n_ops);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < a_vec; i += 128) A[i] = a_img[i];
    for (int i = tid; i < b_vec; i += 128) B[i] = b_img[i];
    for (int i = tid; i < n_ops; i += 128) T[i] = ops[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;

    const uint64_t a_desc = a_desc_hi | (uint64_t)((smem_u32(A) >> 4) & 0x3fff);
    const uint64_t b_desc = b_desc_hi | (uint64_t)((smem_u32(B) >> 4) & 0x3fff);
    const long long t0 = clock64();
    uint32_t parity = 0;
    for (int r = 0; r < reps; ++r) {
        if (tid == 0) {
            auto issue = [&](const MmaOp& op) {
                const uint64_t ad = a_desc + op.a_off16, bd = b_desc + op.b_off16;
                const uint32_t acc = op.d_col >> 31, dcol = tmem + (op.d_col & 0xffffu);
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(dcol),
                    "l"(ad), "l"(bd), "r"(op.idesc), "r"(acc)
                    : "memory");
            };
            int i = 0;
            for (; i + 8 <= n_ops; i += 8) {                 // operands of 8 MMAs fetched ahead of their issue
                MmaOp o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = T[i + j];
#pragma unroll
                for (int j = 0; j < 8; ++j) issue(o[j]);
            }
            for (; i < n_ops; ++i) issue(T[i]);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
        }
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                : "=r"(done)
                : "r"(smem_u32(bar)), "r"(parity)
                : "memory");
        }
        parity ^= 1;
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0 && cycles) cycles[blockIdx.x] = t1 - t0;

    if (blockIdx.x == 0 && d_out) {
        for (int c0 = 0; c0 < out_cols; c0 += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem + ((uint32_t)(32 * warp) << 16) + c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                  "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                  "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                  "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) d_out[(size_t)(32 * warp + lane) * out_cols + c0 + j] = __uint_as_float(v[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols));
}

}  // namespace

// All pointers are device pointers.  a_desc_hi / b_desc_hi: descriptor without the start-address
// field (LBO, SBO, version, layout).  Returns a cudaError_t value.
extern "C" int umma_probe_run(const void* a_img, int a_bytes, const void* b_img, int b_bytes, const void* ops, int n_ops,
                              unsigned long long a_desc_hi, unsigned long long b_desc_hi, int reps, int tmem_cols,
                              int grid, float* d_out, int out_cols, long long* cycles) {
    const size_t smem = (size_t)a_bytes + b_bytes + (size_t)n_ops * sizeof(MmaOp) + 64;
    cudaError_t e = cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    probe_kernel<<<grid, 128, smem>>>((const uint4*)a_img, a_bytes / 16, (const uint4*)b_img, b_bytes / 16,
                                      (const MmaOp*)ops, n_ops, a_desc_hi, b_desc_hi, reps, tmem_cols, d_out, out_cols, cycles);
    return (int)cudaGetLastError();
}
