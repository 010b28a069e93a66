// fic_search_umma.cu -- K2+K3: the full-pool range x domain search as an exact integer
// contraction on the 5th-generation tensor cores (tcgen05, kind::i8, accumulators in
// TMEM), with the scoring / argmin epilogue fused so that no score matrix is ever
// written to HBM.  sm_100a only.
//
// What is computed.  For range block i and domain block j the reference scores
//     error = vR^2 * (1 - (kov / (vR * sqrt(varD)))^2)          (FC:677-683)
// with the exact integers kov = sum (r - rmean)(d - dmean), vR = sum (r - rmean) and
// varD = sum (d - dmean)^2, and keeps the first index with the smallest float error
// (strict <, ascending loop, FC:619-632).  error is a non-increasing function of
// x = |kov| / sqrt(varD) through every rounding step, so the reference's winner is the
// lowest index at which the running maximum of x is (re)attained up to rounding.  The
// epilogue therefore
//   1. gets kov[i][j] exactly from the tensor cores,
//   2. filters with the cheap binary32 score f = |kov| * rsqrt(varD_j) against the
//      running maximum of the row (relative slack 2^-20, far above f's rounding error),
//   3. evaluates the reference's own float/double error formula only for candidates
//      that pass, and applies the reference's strict-< update in ascending index order.
// Result: the same winner index as the reference for every row, bit for bit.
//
// How kov becomes one u8 x s8 GEMM.  With dt = d - dmean_j (|dt| <= 254 for B <= 8),
// split dt = h + l, h = dt >> 1, l = dt - h (both fit s8).  Then
//     kov = sum_k r_k * h_k + sum_k r_k * l_k + rmean_i * (-alpha_j),   alpha_j = sum d - n*dmean_j
// i.e. A row = [ r | r | rmean 0.. ] (u8) and B row = [ h | l | -alpha 0.. ] (s8), K
// padded to a multiple of 32 (one kind::i8 MMA consumes K = 32).  The duplicated `r`
// half of A is not stored twice: the MMA issuer simply points the A descriptor of
// K-slices 2,3 back at slices 0,1 (B = 8).
//
// Data movement.  Operands are packed once per encode by two HBM-bound kernels into
// "blobs" that are already in the canonical no-swizzle K-major UMMA shared-memory
// layout (8x16-byte core matrices, LBO = 128 B between K-adjacent core matrices,
// SBO = KS*256 B between 8-row groups).  A blob is moved with ONE 1-D TMA bulk copy
// (cp.async.bulk, SASS UBLKCP) that completes on an mbarrier; a domain tile blob also
// carries the 128 per-column scales rsqrt(varD) (f32) and sqrt(varD) (f64).
//
// CTA organisation (1 CTA / SM, 640 threads, persistent over work units):
//   warp 0      TMA producer: A super-block (512 range rows, resident for the whole
//               unit) + a ring of domain tiles (128 domains each)
//   warp 1      MMA issuer: per domain tile 4 accumulators (128 rows x 128 domains s32,
//               4 x 128 = all 512 TMEM columns) x NS K-slices of tcgen05.mma
//   warp 2      TMEM allocator
//   warps 4-19  epilogue: warp e owns accumulator q = e / 4, TMEM lanes 32*(e % 4)..+31;
//               one thread owns one range row for the whole unit
// A work unit is (super-block of 512 rows) x (1/n_chunks of the domain tiles); per-unit
// row winners go to part_err / part_idx and are merged in ascending chunk order.
#include "fic_device.cuh"

namespace fic {

namespace {

constexpr int kTileN = 128;        // domains per B tile == MMA N
constexpr int kBlockM = 128;       // rows per accumulator == MMA M
constexpr int kAccs = 4;           // accumulators per tile (TMEM: 4 x 128 columns)
constexpr int kRowsPerSB = kBlockM * kAccs;
constexpr int kEpiWarps = 16;
constexpr int kThreads = (4 + kEpiWarps) * 32;
constexpr int kScaleBytes = kTileN * 4;  // per-column filter scale 1/sqrt(varD), f32

template <int B>
struct Cfg;
template <>
struct Cfg<8> {
    static constexpr int n = 64;
    static constexpr int KS_A = 3;   // physical A K-slices (32 B each): r[0:32] r[32:64] [rmean 0..]
    static constexpr int KS_B = 5;   // h[0:32] h[32:64] l[0:32] l[32:64] [-alpha 0..]
    static constexpr int NS = 5;     // MMA K-slices
    static constexpr int NSTAGE = 7;
    __host__ __device__ static constexpr int amap(int s) { return s < 4 ? (s & 1) : 2; }
};
template <>
struct Cfg<4> {
    static constexpr int n = 16;
    static constexpr int KS_A = 2;   // [r r] [rmean 0..]
    static constexpr int KS_B = 2;   // [h l] [-alpha 0..]
    static constexpr int NS = 2;
    static constexpr int NSTAGE = 8;
    __host__ __device__ static constexpr int amap(int s) { return s; }
};

template <int B>
struct Lay {
    using C = Cfg<B>;
    static constexpr int SBO_A = C::KS_A * 256;
    static constexpr int SBO_B = C::KS_B * 256;
    static constexpr int A_BLOCK_BYTES = (kBlockM / 8) * SBO_A;
    static constexpr int A_SB_BYTES = kAccs * A_BLOCK_BYTES;
    static constexpr int B_OP_BYTES = (kTileN / 8) * SBO_B;
    static constexpr int B_TILE_BYTES = B_OP_BYTES + kScaleBytes;
    static constexpr int SMEM_BYTES = A_SB_BYTES + C::NSTAGE * B_TILE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

// ---------------------------------------------------------------- operand packing ----

// The domain sweep order is a fixed pseudo-random permutation of the pool: chunk P (32 positions)
// of the sweep holds domain chunk (P * mult) mod NCHpad (mult coprime to NCHpad; domains >= ND are
// padding).
// Spatially neighbouring domains have similar scores; sweeping them in raster order would make
// the running maximum of a row climb in long monotone runs (hundreds of "records" per row),
// while a scattered order gives the O(log N) records of an i.i.d. sequence.
// The permutation acts on whole 32-domain chunks (the unit the epilogue flags), so the 32 lanes
// of a refine warp still read 32 adjacent, overlapping domain blocks.
__host__ __device__ __forceinline__ int64_t pos_to_domain(int64_t pos, uint32_t mult, int64_t nchpad)
{
    const uint64_t chunk = (uint64_t)pos >> 5;
    return (int64_t)(((chunk * (uint64_t)mult) % (uint64_t)nchpad) * 32 + ((uint64_t)pos & 31));
}

__device__ __forceinline__ uint32_t pack4(int a, int b, int c, int d)
{
    return (uint32_t)(a & 0xff) | ((uint32_t)(b & 0xff) << 8) | ((uint32_t)(c & 0xff) << 16) | ((uint32_t)(d & 0xff) << 24);
}

// One thread per (padded) domain: centre by the integer mean, split into two s8 digits,
// write the tile blob row and the two per-column scales.
template <int B>
__global__ void __launch_bounds__(128)
k_umma_pack_domains(const uint8_t *__restrict__ dec, const int32_t *__restrict__ dsum,
                    const int32_t *__restrict__ dsq, uint8_t *__restrict__ opB, Geom g, int64_t ntiles,
                    uint32_t mult)
{
    using L = Lay<B>;
    constexpr int n = Cfg<B>::n;
    int64_t pos = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // position in the (permuted) sweep order
    if (pos >= ntiles * kTileN) return;
    int64_t tile = pos / kTileN;
    int row = (int)(pos % kTileN);
    const int64_t j = pos_to_domain(pos, mult, ntiles * (kTileN / 32));
    uint8_t *blob = opB + tile * L::B_TILE_BYTES;
    uint8_t *rowp = blob + (row >> 3) * L::SBO_B + (row & 7) * 16;
    float *rsd = (float *)(blob + L::B_OP_BYTES);
    constexpr int NCH = Cfg<B>::KS_B * 2;  // 16-byte chunks per row
    if (j >= g.ND) {
#pragma unroll
        for (int c = 0; c < NCH; c++) *(uint4 *)(rowp + c * 128) = make_uint4(0, 0, 0, 0);
        rsd[row] = 0.0f;
        return;
    }
    int gx = (int)(j % g.dpw), gy = (int)(j / g.dpw);
    const uint8_t *p = dec + (int64_t)(gy * g.step) * g.sw + gx * g.step;
    int dmean;
    int varD = dom_var(dsum[j], dsq[j], n, &dmean);
    int alpha = dsum[j] - n * dmean;
    constexpr int PCH = n / 16;  // pixel chunks per digit
#pragma unroll
    for (int c = 0; c < PCH; c++) {
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int w = 0; w < 4; w++) {
            int hv[4], lv[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                int k = c * 16 + w * 4 + e;
                int dt = (int)__ldg(p + (int64_t)(k / B) * g.sw + (k % B)) - dmean;
                hv[e] = dt >> 1;
                lv[e] = dt - hv[e];
            }
            hw[w] = pack4(hv[0], hv[1], hv[2], hv[3]);
            lw[w] = pack4(lv[0], lv[1], lv[2], lv[3]);
        }
        *(uint4 *)(rowp + c * 128) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        *(uint4 *)(rowp + (PCH + c) * 128) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    }
    *(uint4 *)(rowp + (2 * PCH) * 128) = make_uint4((uint32_t)((-alpha) & 0xff), 0, 0, 0);
    *(uint4 *)(rowp + (2 * PCH + 1) * 128) = make_uint4(0, 0, 0, 0);
    rsd[row] = varD > 0 ? __double2float_rn(__ddiv_rn(1.0, __dsqrt_rn((double)varD))) : 0.0f;
}

// One thread per (padded) range row of the slice [j0, j1): raw pixels + integer mean.
template <int B>
__global__ void __launch_bounds__(128)
k_umma_pack_ranges(const uint8_t *__restrict__ src, const int32_t *__restrict__ rsum, uint8_t *__restrict__ opA,
                   int32_t *__restrict__ vRout, Geom g, int64_t j0, int64_t j1, int64_t rows_padded)
{
    using L = Lay<B>;
    constexpr int n = Cfg<B>::n;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows_padded) return;
    int64_t sb = i / kRowsPerSB;
    int rr = (int)(i % kRowsPerSB);
    int blk = rr / kBlockM, row = rr % kBlockM;
    uint8_t *rowp = opA + sb * L::A_SB_BYTES + blk * L::A_BLOCK_BYTES + (row >> 3) * L::SBO_A + (row & 7) * 16;
    constexpr int NCH = Cfg<B>::KS_A * 2;
    int64_t j = j0 + i;
    if (j >= j1) {
#pragma unroll
        for (int c = 0; c < NCH; c++) *(uint4 *)(rowp + c * 128) = make_uint4(0, 0, 0, 0);
        vRout[i] = 0;
        return;
    }
    int xr = (int)(j % g.rpw), yr = (int)(j / g.rpw);
    const uint8_t *p = src + (int64_t)(yr * B) * g.W + xr * B;
    int rs = rsum[j];
    int rmean = rs / n;
    vRout[i] = rs - n * rmean;
    constexpr int PCH = n / 16;
#pragma unroll
    for (int c = 0; c < PCH; c++) {
        uint32_t w4[4];
#pragma unroll
        for (int w = 0; w < 4; w++) {
            int k = c * 16 + w * 4;
            // 4 consecutive k share a pixel row for B >= 4; 4-byte aligned since xr*B, k%B are multiples of 4
            w4[w] = *(const uint32_t *)(p + (int64_t)(k / B) * g.W + (k % B));
        }
        uint4 v = make_uint4(w4[0], w4[1], w4[2], w4[3]);
        *(uint4 *)(rowp + c * 128) = v;
        if (B == 4) *(uint4 *)(rowp + (PCH + c) * 128) = v;  // [r r] shares one 32-byte slice
    }
    constexpr int XCH = (B == 4) ? 2 * PCH : PCH;
    *(uint4 *)(rowp + XCH * 128) = make_uint4((uint32_t)rmean, 0, 0, 0);
    *(uint4 *)(rowp + (XCH + 1) * 128) = make_uint4(0, 0, 0, 0);
}

// ---------------------------------------------------------------- PTX wrappers -------

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok;
}
// Bounded wait: a pipeline bug must end in an error, never in a hung GPU.  status is
// pinned host memory (mapped), so the code survives the trap.
__device__ __noinline__ void mbar_timeout(volatile int *status, int code)
{
    if (status) {
        *status = code;
        __threadfence_system();
    }
    __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, volatile int *status, int code)
{
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 2000000000ll) mbar_timeout(status, code);
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ uint32_t elect_one()
{
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 r;\n\t"
        "elect.sync r|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
// Accumulators are pre-loaded with the bit pattern of 1.5 * 2^23 ("magic" bias): after the integer
// MMA, lane bits = 0x4B400000 + kov, which read as binary32 is exactly 12582912 + kov for
// |kov| < 2^22 (B <= 8: |kov| <= 64*255*254).  One FADD then yields float(kov) on the FMA pipe
// instead of an I2FP on the half-rate ALU pipe.
constexpr uint32_t kMagicBits = 0x4B400000u;
constexpr float kMagic = 12582912.0f;

__device__ __forceinline__ void tmem_st8_const(uint32_t taddr, uint32_t c)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(c)
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// (a0, a1) = ((a0, a1) - magic) * (s0, s1) with the packed binary32 pipe (FADD2 / FMUL2).
__device__ __forceinline__ void unbias_scale2(uint32_t a0, uint32_t a1, float s0, float s1, float &f0, float &f1)
{
    unsigned long long v, sc, mg, x, f;
    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "r"(a0), "r"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(sc) : "f"(s0), "f"(s1));
    asm("mov.b64 %0, {%1, %1};" : "=l"(mg) : "f"(-kMagic));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(x) : "l"(v), "l"(mg));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(f) : "l"(x), "l"(sc));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(f0), "=f"(f1) : "l"(f));
}

// (f0, f1) = (x0, x1) * (s0, s1), packed binary32 multiply (FMUL2).
__device__ __forceinline__ void scale2(float x0, float x1, float s0, float s1, float &f0, float &f1)
{
    unsigned long long x, sc, f;
    asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(x0), "f"(x1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(sc) : "f"(s0), "f"(s1));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(f) : "l"(x), "l"(sc));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(f0), "=f"(f1) : "l"(f));
}

__device__ __forceinline__ float4 lds_f4(uint32_t saddr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, no-swizzle shared-memory matrix descriptor (PTX "matrix descriptor";
// cute::UMMA::SmemDescriptor): start address, leading (K) and stride (M/N) byte offsets
// in 16-byte units, descriptor version 1 (Blackwell) in bits 46-47, layout type 0.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}

// kind::i8 instruction descriptor (cute::UMMA::InstrDescriptor): D = s32 (c_format 2 at
// bit 4), A = u8 (0 at bit 7), B = s8 (1 at bit 10), both K-major, N>>3 at bit 17, M>>4
// at bit 24.
constexpr uint32_t kIdesc = (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(kTileN >> 3) << 17) |
                            ((uint32_t)(kBlockM >> 4) << 24);

// ---------------------------------------------------------------- epilogue / refine --

constexpr int kDefaultEpi = 0;
constexpr int kFlagCap = 32;                            // flagged 32-column chunks kept per (row, unit)
constexpr float kOneMinusEps = 1.0f - 1.9073486328125e-06f;  // 1 - 2^-19 (applied to the squared score)

// ---------------------------------------------------------------- the search kernel --

// DBG (probe builds only): 1 = skip the scoring math, 3 = skip the TMEM loads too.  DUMP: write every
// accumulator to `dump` (probe's exactness check).  The product runs <B, 0, false>.
// EPI: 0 = I2FP + FMUL2 scoring on plain s32 accumulators; 1 = magic-biased accumulators (FADD2 + FMUL2,
// tcgen05.st re-bias).
template <int B, int DBG, bool DUMP, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
k_umma_search(const uint8_t *__restrict__ opA, const uint8_t *__restrict__ opB, const int32_t *__restrict__ vRarr,
              int32_t *__restrict__ flag_list, int32_t *__restrict__ flag_cnt, int n_sb, int n_chunks, int ntiles,
              int64_t rows_padded, int32_t *__restrict__ dump, int64_t dump_ld, volatile int *status,
              uint32_t lbo_bytes_a, uint32_t sbo_bytes_a, uint32_t lbo_bytes_b, uint32_t sbo_bytes_b)
{
    using C = Cfg<B>;
    using L = Lay<B>;
    constexpr int NSTAGE = C::NSTAGE;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [A super-block][NSTAGE domain tiles][barriers]
    uint8_t *smem = (uint8_t *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;
    uint8_t *sB = smem + L::A_SB_BYTES;
    uint64_t *bars = (uint64_t *)(sB + NSTAGE * L::B_TILE_BYTES);
    // barrier map
    const uint32_t bar0 = smem_u32(bars);
    auto BAR_B_FULL = [&](int s) { return bar0 + 8u * s; };
    auto BAR_B_EMPTY = [&](int s) { return bar0 + 8u * (NSTAGE + s); };
    auto BAR_T_FULL = [&](int q) { return bar0 + 8u * (2 * NSTAGE + q); };
    auto BAR_T_EMPTY = [&](int q) { return bar0 + 8u * (2 * NSTAGE + kAccs + q); };
    const uint32_t BAR_A_FULL = bar0 + 8u * (2 * NSTAGE + 2 * kAccs);
    const uint32_t BAR_A_EMPTY = BAR_A_FULL + 8u;
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * NSTAGE + 2 * kAccs + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < NSTAGE; s++) {
            mbar_init(BAR_B_FULL(s), 1);
            mbar_init(BAR_B_EMPTY(s), kEpiWarps);
        }
        for (int q = 0; q < kAccs; q++) {
            mbar_init(BAR_T_FULL(q), 1);
            mbar_init(BAR_T_EMPTY(q), 2 * kEpiWarps / kAccs);  // 4 lane quarters x 2 column halves
        }
        mbar_init(BAR_A_FULL, 1);
        mbar_init(BAR_A_EMPTY, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_units = n_sb * n_chunks;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, a_phase = 0;
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                int sb = u / n_chunks, ch = u % n_chunks;
                int t0 = (int)((int64_t)ch * ntiles / n_chunks), t1 = (int)((int64_t)(ch + 1) * ntiles / n_chunks);
                mbar_wait(BAR_A_EMPTY, a_phase ^ 1, status, 1);
                mbar_expect_tx(BAR_A_FULL, L::A_SB_BYTES);
                bulk_g2s(smem_u32(sA), opA + (int64_t)sb * L::A_SB_BYTES, L::A_SB_BYTES, BAR_A_FULL);
                for (int t = t0; t < t1; t++) {
                    mbar_wait(BAR_B_EMPTY(stage), phase ^ 1, status, 2);
                    mbar_expect_tx(BAR_B_FULL(stage), L::B_TILE_BYTES);
                    bulk_g2s(smem_u32(sB + stage * L::B_TILE_BYTES), opB + (int64_t)t * L::B_TILE_BYTES,
                             L::B_TILE_BYTES, BAR_B_FULL(stage));
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
                a_phase ^= 1;
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // The whole warp walks the (warp-uniform) loop so that addresses and descriptors live in
        // uniform registers; one elected lane issues tcgen05.mma / tcgen05.commit.
        uint32_t stage = 0, phase = 0, a_phase = 0, t_phase = 0;  // t_phase: bit q
        const uint32_t elected = elect_one();
        // descriptors differ only in the 14-bit start-address field (16-byte units)
        const uint64_t a_desc0 = make_desc(smem_u32(sA), lbo_bytes_a, sbo_bytes_a);
        const uint64_t b_desc0 = make_desc(smem_u32(sB), lbo_bytes_b, sbo_bytes_b);
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            int ch = u % n_chunks;
            int t0 = (int)((int64_t)ch * ntiles / n_chunks), t1 = (int)((int64_t)(ch + 1) * ntiles / n_chunks);
            mbar_wait(BAR_A_FULL, a_phase, status, 3);
            for (int t = t0; t < t1; t++) {
                mbar_wait(BAR_B_FULL(stage), phase, status, 4);
                tc_fence_after();
                const uint64_t b_desc = b_desc0 + (uint64_t)((stage * L::B_TILE_BYTES) >> 4);
#pragma unroll
                for (int q = 0; q < kAccs; q++) {
                    // EPI 1: the epilogue (re-)biased accumulator q and arrived, also once before the first
                    // tile; EPI 0: a fresh barrier passes the inverted-parity wait.
                    mbar_wait(BAR_T_EMPTY(q), ((t_phase >> q) & 1) ^ (EPI == 1 ? 0u : 1u), status, 5);
                    tc_fence_after();
                    if (elected) {
#pragma unroll
                        for (int s = 0; s < C::NS; s++) {
                            const uint64_t ad = a_desc0 + (uint64_t)((q * L::A_BLOCK_BYTES + C::amap(s) * 256) >> 4);
                            const uint64_t bd = b_desc + (uint64_t)((s * 256) >> 4);
                            tc_mma_i8(tmem_base + q * kTileN, ad, bd, kIdesc, (EPI == 1 || s > 0) ? 1u : 0u);
                        }
                        tc_commit(BAR_T_FULL(q));
                    }
                    __syncwarp();
                    t_phase ^= 1u << q;
                }
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
            if (elected) tc_commit(BAR_A_EMPTY);  // every MMA that reads this A super-block has completed
            __syncwarp();
            a_phase ^= 1;
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        // Warp e: TMEM lane quarter lq = e % 4 (== warp % 4, the quarter this warp may access), column half
        // `half` (64 of the 128 domains of a tile) of the two accumulators qa and qa + 2.  Two warps share
        // each (accumulator, lane quarter), so an accumulator is drained in two chunk-times, and every warp
        // alternates between two accumulators so that it always has one ready while the other is refilled.
        const int e = warp - 4;
        const int lq = e & 3, kk = e >> 2;
        const int half = kk & 1, qa = kk >> 1;
        const uint32_t t_lane0 = tmem_base + ((uint32_t)(lq * 32) << 16) + half * 64;
        uint32_t stage = 0, phase = 0, tf_phase = 0;
        if (EPI == 1) {
#pragma unroll
            for (int sl = 0; sl < 2; sl++) {
                const uint32_t ta = t_lane0 + (qa + 2 * sl) * kTileN;
#pragma unroll
                for (int c8 = 0; c8 < 8; c8++) tmem_st8_const(ta + c8 * 8, kMagicBits);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(BAR_T_EMPTY(qa));
                mbar_arrive(BAR_T_EMPTY(qa + 2));
            }
        }
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            int sb = u / n_chunks, ch = u % n_chunks;
            int t0 = (int)((int64_t)ch * ntiles / n_chunks), t1 = (int)((int64_t)(ch + 1) * ntiles / n_chunks);
            // Running filter state of this thread's two rows.  thresh: chunks whose best filter score is
            // <= thresh cannot hold the reference's winner or one of its float ties.  vR == 0: every
            // candidate scores error 0 and the first one wins (FC:677-678, FC:627) -> never flag.
            int64_t row[2];
            float thresh[2], fmax[2], tie_abs[2];
            int cnt[2];
            int32_t *my_list[2];
#pragma unroll
            for (int sl = 0; sl < 2; sl++) {
                row[sl] = (int64_t)sb * kRowsPerSB + (qa + 2 * sl) * kBlockM + lq * 32 + lane;
                const int vR = vRarr[row[sl]];
                thresh[sl] = (vR == 0) ? __int_as_float(0x7f800000) : -1.0f;
                fmax[sl] = 0.0f;
                tie_abs[sl] = (float)(vR * vR) * 4.76837158203125e-07f;  // vR^2 * 2^-21
                cnt[sl] = 0;
                my_list[sl] = flag_list + (((int64_t)ch * rows_padded + row[sl]) * 2 + half) * kFlagCap;
            }
            for (int t = t0; t < t1; t++) {
                mbar_wait(BAR_B_FULL(stage), phase, status, 6);
                const uint32_t rsd_s = smem_u32(sB + stage * L::B_TILE_BYTES + L::B_OP_BYTES) + half * 64 * 4;
#pragma unroll
                for (int sl = 0; sl < 2; sl++) {
                    const int q = qa + 2 * sl;
                    const uint32_t ta = t_lane0 + q * kTileN;
                    mbar_wait(BAR_T_FULL(q), tf_phase, status, 7);
                    tc_fence_after();
#pragma unroll
                    for (int cc = 0; cc < 2; cc++) {
                        uint32_t v[32];
                        if (!(DBG & 2)) {
                            tmem_ld32(ta + cc * 32, v);
                            tmem_ld_wait();
                        }
                        if (EPI == 1) {  // re-bias the columns just read
#pragma unroll
                            for (int c8 = 0; c8 < 4; c8++) tmem_st8_const(ta + cc * 32 + c8 * 8, kMagicBits);
                        }
                        if (cc == 1) {
                            // last access to this half of accumulator q for this tile: hand it back
                            if (EPI == 1) tmem_st_wait();
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(BAR_T_EMPTY(q));
                        }
                        if (DBG & 1) continue;  // probe only: measure the pipeline without the scoring math
                        float m = 0.0f;
#pragma unroll
                        for (int k = 0; k < 32; k += 4) {
                            float4 s4 = lds_f4(rsd_s + (cc * 32 + k) * 4);
                            float f0, f1, f2, f3;
                            if (EPI == 1) {
                                unbias_scale2(v[k + 0], v[k + 1], s4.x, s4.y, f0, f1);
                                unbias_scale2(v[k + 2], v[k + 3], s4.z, s4.w, f2, f3);
                            } else {
                                scale2(__int2float_rn((int)v[k + 0]), __int2float_rn((int)v[k + 1]), s4.x, s4.y, f0, f1);
                                scale2(__int2float_rn((int)v[k + 2]), __int2float_rn((int)v[k + 3]), s4.z, s4.w, f2, f3);
                            }
                            m = fmaxf(fmaxf(m, fabsf(f0)), fabsf(f1));
                            m = fmaxf(fmaxf(m, fabsf(f2)), fabsf(f3));
                        }
                        const int c = half * 2 + cc;  // chunk of the tile
                        if (DUMP) {
#pragma unroll
                            for (int k = 0; k < 32; k++)
                                dump[row[sl] * dump_ld + (int64_t)t * kTileN + c * 32 + k] =
                                    (int)(v[k] - (EPI == 1 ? kMagicBits : 0u));
                        }
                        if (m > thresh[sl]) {  // may hold the winner or one of its float ties: exact work is deferred
                            if (cnt[sl] < kFlagCap) my_list[sl][cnt[sl]] = t * (kTileN / 32) + c;
                            cnt[sl]++;
                            fmax[sl] = fmaxf(fmax[sl], m);
                            thresh[sl] = sqrtf(fmaxf(fmax[sl] * fmax[sl] * kOneMinusEps - tie_abs[sl], 0.0f));
                        }
                    }
                }
                // the tile's scales are consumed
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR_B_EMPTY(stage));
                tf_phase ^= 1;
                if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
            }
#pragma unroll
            for (int sl = 0; sl < 2; sl++) flag_cnt[((int64_t)ch * rows_padded + row[sl]) * 2 + half] = cnt[sl];
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// Refine: exact evaluation of every candidate of every flagged chunk with the reference's own
// arithmetic (FC:655-687), one warp per range row, lane = candidate within the chunk; the winner is
// the lexicographic (error, index) minimum = the reference's first index with the smallest error.
// A row whose flag list overflowed (adversarial score order) is rescanned in full -- slow, exact.
template <int B>
__device__ __forceinline__ float refine_eval(const int *s_rt, const uint8_t *__restrict__ dec,
                                             const int32_t *__restrict__ dsum, const int32_t *__restrict__ dsq,
                                             const Geom &g, int vR, int64_t idx)
{
    constexpr int n = B * B;
    int gx = (int)(idx % g.dpw), gy = (int)(idx / g.dpw);
    const uint8_t *p = dec + (int64_t)(gy * g.step) * g.sw + gx * g.step;
    int dot = 0;
#pragma unroll
    for (int ry = 0; ry < B; ry++) {
        const uint8_t *row = p + (int64_t)ry * g.sw;
        if (B == 8) {  // gx*step and sw are even: 2-byte aligned
#pragma unroll
            for (int rx = 0; rx < B; rx += 2) {
                unsigned v = __ldg((const unsigned short *)(row + rx));
                dot += s_rt[ry * B + rx] * (int)(v & 0xff) + s_rt[ry * B + rx + 1] * (int)(v >> 8);
            }
        } else {
#pragma unroll
            for (int rx = 0; rx < B; rx++) dot += s_rt[ry * B + rx] * (int)__ldg(row + rx);
        }
    }
    int dmean;
    int varD = dom_var(dsum[idx], dsq[idx], n, &dmean);
    return grey_error(dot - dmean * vR, vR, __dsqrt_rn((double)varD));
}

template <int B>
__global__ void __launch_bounds__(128)
k_umma_refine(const uint8_t *__restrict__ src, const uint8_t *__restrict__ dec, const int32_t *__restrict__ dsum,
              const int32_t *__restrict__ dsq, const int32_t *__restrict__ rsum,
              const int32_t *__restrict__ flag_list, const int32_t *__restrict__ flag_cnt, int n_chunks,
              int64_t rows_padded, int64_t rows, int32_t *__restrict__ best, Geom g, int64_t j0, uint32_t mult,
              int64_t nchpad)
{
    constexpr int n = B * B;
    __shared__ int s_rt_all[4][n];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * 4 + warp;
    if (i >= rows) return;
    const int64_t j = j0 + i;
    int *s_rt = s_rt_all[warp];
    const int rs = rsum[j];
    const int rmean = rs / n, vR = rs - n * rmean;
    if (vR == 0) {  // FC:677-678 + FC:627: all errors are 0, the first candidate wins
        if (lane == 0) best[j] = 0;
        return;
    }
    const int xr = (int)(j % g.rpw), yr = (int)(j / g.rpw);
    for (int k = lane; k < n; k += 32)
        s_rt[k] = (int)src[(int64_t)(yr * B + k / B) * g.W + xr * B + (k % B)] - rmean;
    __syncwarp();
    float be = 10000000.0f;  // FC:615
    int bi = 0x7fffffff;
    bool overflow = false;
    for (int lh = 0; lh < 2 * n_chunks && !overflow; lh++) {  // (domain chunk of the unit, column half) lists
        const int64_t li = ((int64_t)(lh >> 1) * rows_padded + i) * 2 + (lh & 1);
        const int cnt = flag_cnt[li];
        if (cnt > kFlagCap) { overflow = true; break; }
        const int32_t *lst = flag_list + li * kFlagCap;
        for (int e = 0; e < cnt; e++) {
            const int64_t idx = pos_to_domain((int64_t)lst[e] * 32 + lane, mult, nchpad);
            if (idx < g.ND) {
                float err = refine_eval<B>(s_rt, dec, dsum, dsq, g, vR, idx);
                if (err < be || (err == be && (int)idx < bi)) { be = err; bi = (int)idx; }
            }
        }
    }
    if (overflow) {
        be = 10000000.0f;
        bi = 0x7fffffff;
        for (int64_t idx = lane; idx < g.ND; idx += 32) {
            float err = refine_eval<B>(s_rt, dec, dsum, dsq, g, vR, idx);
            if (err < be) { be = err; bi = (int)idx; }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        float e2 = __shfl_down_sync(0xffffffffu, be, o);
        int i2 = __shfl_down_sync(0xffffffffu, bi, o);
        if (e2 < be || (e2 == be && i2 < bi)) { be = e2; bi = i2; }
    }
    if (lane == 0) best[j] = bi == 0x7fffffff ? 0 : bi;
}

inline int64_t pad_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

struct Plan {
    int64_t rp;     // rows padded to whole super-blocks
    int n_sb, ntiles, n_chunks;
    uint32_t mult;  // sweep-order multiplier (see pos_to_domain)
};

inline uint64_t gcd_u64(uint64_t a, uint64_t b) { while (b) { uint64_t t = a % b; a = b; b = t; } return a; }

// Split the domain sweep so that the unit count fills whole waves of num_sms CTAs.
inline Plan make_plan(const Geom &g, int64_t rows, int num_sms)
{
    Plan p;
    p.rp = pad_up(rows, kRowsPerSB);
    p.n_sb = (int)(p.rp / kRowsPerSB);
    p.ntiles = (int)((g.ND + kTileN - 1) / kTileN);
    p.n_chunks = 1;
    double best_eff = 0;
    for (int c = 1; c <= 8 && c <= p.ntiles; c++) {
        int64_t units = (int64_t)p.n_sb * c;
        int64_t waves = (units + num_sms - 1) / num_sms;
        double eff = (double)units / (double)(waves * num_sms);
        if (eff > best_eff + 0.02) { best_eff = eff; p.n_chunks = c; }
    }
    const uint64_t nchpad = (uint64_t)p.ntiles * (kTileN / 32);
    uint64_t m = (uint64_t)((double)nchpad * 0.6180339887498949) | 1u;  // golden-ratio stride, odd
    while (gcd_u64(m, nchpad) != 1) m += 2;
    p.mult = nchpad <= 8 ? 1u : (uint32_t)(m % nchpad);
    return p;
}

template <int B>
size_t opA_bytes_t(const Geom &g, int64_t rows, int num_sms)
{
    Plan p = make_plan(g, rows, num_sms);
    // [A blobs][vR s32][flag_cnt s32 x n_chunks x 2][flag_list s32 x n_chunks x 2 x kFlagCap]
    return (size_t)p.n_sb * Lay<B>::A_SB_BYTES + (size_t)p.rp * 4 + (size_t)p.rp * p.n_chunks * 2 * 4 * (1 + kFlagCap) + 1024;
}

template <int B>
int launch_t(const Work &w, const Geom &g, int64_t j0, int64_t j1, int num_sms, cudaStream_t s, const char **err,
             int32_t *dump, int64_t dump_ld, int *status_dev, int variant, cudaEvent_t k0 = nullptr,
             cudaEvent_t k1 = nullptr, uint32_t dbg = 0)
{
    using L = Lay<B>;
    int64_t rows = j1 - j0;
    if (rows <= 0) return 0;
    Plan p = make_plan(g, rows, num_sms);
    const int64_t rp = p.rp;
    uint8_t *opA = w.opA;
    int32_t *vR = (int32_t *)(opA + (size_t)p.n_sb * L::A_SB_BYTES);
    int32_t *flag_cnt = vR + rp;
    int32_t *flag_list = flag_cnt + rp * p.n_chunks * 2;
    int launches = 0;
    k_umma_pack_domains<B><<<(unsigned)(((int64_t)p.ntiles * kTileN + 127) / 128), 128, 0, s>>>(w.dec, w.dsum, w.dsq, w.opB, g, p.ntiles, p.mult);
    k_umma_pack_ranges<B><<<(unsigned)((rp + 127) / 128), 128, 0, s>>>(w.src, w.rsum, opA, vR, g, j0, j1, rp);
    launches += 2;
    using KernelT = void (*)(const uint8_t *, const uint8_t *, const int32_t *, int32_t *, int32_t *, int, int, int,
                             int64_t, int32_t *, int64_t, volatile int *, uint32_t, uint32_t, uint32_t, uint32_t);
    // dbg (probe only): bit 3 selects the magic-bias epilogue; low bits 1 / 3 strip the scoring math / TMEM loads
    KernelT kern = k_umma_search<B, 0, false, kDefaultEpi>;
    const int epi = (dbg & 8u) ? 1 : ((dbg & 16u) ? 0 : kDefaultEpi);
    const int strip = (int)(dbg & 3u);
    if (dump) kern = epi ? k_umma_search<B, 0, true, 1> : k_umma_search<B, 0, true, 0>;
    else if (strip == 1) kern = epi ? k_umma_search<B, 1, false, 1> : k_umma_search<B, 1, false, 0>;
    else if (strip == 3) kern = epi ? k_umma_search<B, 3, false, 1> : k_umma_search<B, 3, false, 0>;
    else kern = epi ? k_umma_search<B, 0, false, 1> : k_umma_search<B, 0, false, 0>;
    cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::SMEM_BYTES);
    if (ce != cudaSuccess) { *err = cudaGetErrorString(ce); return -1; }
    int n_units = p.n_sb * p.n_chunks;
    int grid = n_units < num_sms ? n_units : num_sms;
    uint32_t lbo_a = 128, sbo_a = L::SBO_A, lbo_b = 128, sbo_b = L::SBO_B;
    if (variant == 1) { lbo_a = L::SBO_A; sbo_a = 128; lbo_b = L::SBO_B; sbo_b = 128; }  // probe only
    if (k0) cudaEventRecord(k0, s);
    kern<<<grid, kThreads, L::SMEM_BYTES, s>>>(opA, w.opB, vR, flag_list, flag_cnt, p.n_sb, p.n_chunks, p.ntiles, rp,
                                               dump, dump_ld, status_dev, lbo_a, sbo_a, lbo_b, sbo_b);
    if (k1) cudaEventRecord(k1, s);
    k_umma_refine<B><<<(unsigned)((rows + 3) / 4), 128, 0, s>>>(w.src, w.dec, w.dsum, w.dsq, w.rsum, flag_list, flag_cnt,
                                                                p.n_chunks, rp, rows, w.best, g, j0, p.mult,
                                                                (int64_t)p.ntiles * (kTileN / 32));
    launches += 2;
    ce = cudaGetLastError();
    if (ce != cudaSuccess) { *err = cudaGetErrorString(ce); return -1; }
    return launches;
}

}  // namespace

void umma_sweep_order(const Geom &g, int64_t rows, int num_sms, uint32_t *mult, int64_t *nchpad)
{
    Plan p = make_plan(g, rows, num_sms);
    *mult = p.mult;
    *nchpad = (int64_t)p.ntiles * (kTileN / 32);
}

bool umma_applicable(const Geom &g)
{
    return g.C == 1 && (g.B == 8 || g.B == 4) && g.wk == g.dpw && g.wk == g.dph;
}

size_t umma_opA_bytes(const Geom &g, int64_t j0, int64_t j1, int num_sms)
{
    return g.B == 8 ? opA_bytes_t<8>(g, j1 - j0, num_sms) : opA_bytes_t<4>(g, j1 - j0, num_sms);
}

size_t umma_opB_bytes(const Geom &g)
{
    int64_t ntiles = (g.ND + kTileN - 1) / kTileN;
    return (size_t)ntiles * (g.B == 8 ? Lay<8>::B_TILE_BYTES : Lay<4>::B_TILE_BYTES);
}

int launch_search_umma(const Work &w, const Geom &g, int64_t j0, int64_t j1, int num_sms, cudaStream_t s,
                       const char **err, cudaEvent_t k0, cudaEvent_t k1)
{
    if (g.B == 8) return launch_t<8>(w, g, j0, j1, num_sms, s, err, nullptr, 0, nullptr, 0, k0, k1);
    return launch_t<4>(w, g, j0, j1, num_sms, s, err, nullptr, 0, nullptr, 0, k0, k1);
}

// Debug entry used by tools/umma_probe: also dumps the raw accumulators (kov) of every
// (row, domain) pair, and lets the probe pick the descriptor variant.
int launch_search_umma_debug(const Work &w, const Geom &g, int64_t j0, int64_t j1, int num_sms, cudaStream_t s,
                             const char **err, int32_t *dump, int64_t dump_ld, int *status_dev, int variant,
                             uint32_t dbg, cudaEvent_t k0, cudaEvent_t k1)
{
    if (g.B == 8) return launch_t<8>(w, g, j0, j1, num_sms, s, err, dump, dump_ld, status_dev, variant, k0, k1, dbg);
    return launch_t<4>(w, g, j0, j1, num_sms, s, err, dump, dump_ld, status_dev, variant, k0, k1, dbg);
}

}  // namespace fic
