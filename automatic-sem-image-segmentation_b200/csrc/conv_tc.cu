// Implicit-GEMM convolutions on the 5th-generation tensor cores (tcgen05.mma, accumulator in TMEM).
//
// im2col-free: a CTA stages the NHWC halo of its 16x8 output-pixel tile ONCE into shared memory in
// "channel-group planes" [Cin/8][halo_h][halo_w][8 x bf16], which is exactly the no-swizzle K-major UMMA
// canonical layout (8 consecutive pixels of one row = one 128-byte core matrix, SBO = halo row pitch,
// LBO = plane pitch).  The 3x3 taps are then nine shared-memory DESCRIPTORS that point at shifted starts
// inside the same halo tile -- no im2col matrix, no re-load per tap.  Weights are pre-packed on the device
// into the matching [tap][Cin/8][Cout][8] image.  One elected thread issues the MMAs; all four warps drain
// the 128-lane accumulator with tcgen05.ld in the epilogue (bias, fp32 BatchNorm/InstanceNorm moments,
// bf16 store at a channel offset of the destination buffer = skip-concat for free).
//
// Replaces F.conv2d / its data gradient under keras.layers.Conv2D for the stride-1 3x3 and 1x1 layers
// (UNet_Segmentation.py:421,465-468,490-499; CycleGAN.py:327,333).
#include "common.cuh"

namespace semb {

constexpr int TILE_H = 16, TILE_W = 8;          // 128 output pixels = UMMA M
constexpr int TC_THREADS = 128;

struct TcPlan { int KC, NC, nchunks, kchunks, tmem_cols; };

// Shared-memory budget: A planes + B taps must leave room for 2 CTAs per SM.
static inline TcPlan tc_plan(int Cin, int Cout, int taps) {
    TcPlan p;
    const int c16 = (Cout + 15) / 16 * 16;
    p.nchunks = (c16 + 255) / 256;
    p.NC = ((c16 + p.nchunks - 1) / p.nchunks + 15) / 16 * 16;
    const int cin16 = (Cin + 15) / 16 * 16;
    int kc = 64;
    while (kc > 16 && (size_t)taps * kc * p.NC * 2 + (size_t)kc * 362 > 96 * 1024) kc >>= 1;
    if (kc > cin16) kc = cin16 <= 16 ? 16 : (cin16 <= 32 ? 32 : 64);
    p.KC = kc;
    p.kchunks = (Cin + kc - 1) / kc;
    p.tmem_cols = p.NC <= 32 ? 32 : (p.NC <= 64 ? 64 : (p.NC <= 128 ? 128 : 256));
    return p;
}

// ---- weight packing ----------------------------------------------------------------------------------
// dst[nchunk][kchunk][tap][KC/8][NC][8] (bf16)  <-  w[r][s][ci][co] (fp32 HWIO), zero padded.
// flip: the packed operator is the stride-1 data gradient: k runs over co, n over ci, taps mirrored.
__global__ void pack_weights_kernel(const float* __restrict__ w, int R, int S, int Cin, int Cout, int flip, TcPlan p,
                                    bf16* __restrict__ dst, long long total) {
    const int taps = R * S;
    const int K = flip ? Cout : Cin;      // reduction channels of the packed operator
    const int N = flip ? Cin : Cout;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long t = i;
        const int ke = (int)(t % 8); t /= 8;
        const int nl = (int)(t % p.NC); t /= p.NC;
        const int k8 = (int)(t % (p.KC / 8)); t /= (p.KC / 8);
        const int tap = (int)(t % taps); t /= taps;
        const int kch = (int)(t % p.kchunks); t /= p.kchunks;
        const int nch = (int)t;
        const int k = kch * p.KC + k8 * 8 + ke, n = nch * p.NC + nl;
        float v = 0.f;
        if (k < K && n < N) {
            if (!flip) v = w[((size_t)tap * Cin + k) * Cout + n];
            else v = w[((size_t)(taps - 1 - tap) * Cin + n) * Cout + k];
        }
        dst[i] = __float2bfloat16_rn(v);
    }
}

// ---- PTX wrappers ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    // bounded spin: a descriptor bug must trap, not hang the GPU box
    const long long t0 = clock64();
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(COLS) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, majors, N>>3 at bit 17, M>>4 at bit 24
__device__ __forceinline__ uint32_t instr_desc(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct TcArgs {
    int N, H, W, OH, OW, Cin, Cout, R, S, pad_t, pad_l, pad_mode;
    const bf16* x; int x_pitch, x_coff;
    bf16* y; int y_pitch, y_coff;
    const bf16* wp; const float* bias;
    double* stats; int stats_nstride, stats_cstride;
    int accumulate;
    int tiles_x, tiles_y;
    TcPlan p;
    int plane_bytes, halo_h, halo_w;
};

template <int COLS>
__global__ void __launch_bounds__(TC_THREADS) conv_tc_kernel(const TcArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int taps = a.R * a.S;
    const int KC = a.p.KC, NC = a.p.NC;
    uint8_t* As = smem;                                              // [KC/8][plane]
    uint8_t* Bs = smem + (KC / 8) * a.plane_bytes;                   // [tap][KC/8][NC][16 B]
    float* part = reinterpret_cast<float*>(Bs + (size_t)taps * KC * NC * 2);   // [2][4][NC] epilogue partials
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x;
    const int n = tile / (a.tiles_x * a.tiles_y);
    const int trem = tile % (a.tiles_x * a.tiles_y);
    const int y0 = (trem / a.tiles_x) * TILE_H, x0 = (trem % a.tiles_x) * TILE_W;
    const int nchunk = blockIdx.y;
    const int halo_pix = a.halo_h * a.halo_w;

    if (warp == 0) tmem_alloc<COLS>(smem_u32(&tmem_slot));
    if (tid == 0) mbar_init(smem_u32(&mbar), 1);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = instr_desc(128, NC, 0, 0);
    uint32_t phase = 0;

    for (int kc = 0; kc < a.p.kchunks; ++kc) {
        const int cvalid = min(KC, a.Cin - kc * KC);                // multiple of 8
        const int ksteps = (cvalid + 15) / 16;
        const int planes = 2 * ksteps;
        // ---- A: halo tile of this channel chunk -> channel-group planes
        for (int idx = tid; idx < planes * halo_pix; idx += TC_THREADS) {
            const int k8 = idx % planes, pix = idx / planes;
            const int hy = pix / a.halo_w, hx = pix % a.halo_w;
            int iy = y0 - a.pad_t + hy, ix = x0 - a.pad_l + hx;
            if (a.pad_mode == SEMB_PAD_REFLECT) {
                // positions that only feed out-of-range output pixels may fall outside the reflectable band
                if (iy > -a.H && iy < 2 * a.H - 1) iy = reflect_index(iy, a.H);
                if (ix > -a.W && ix < 2 * a.W - 1) ix = reflect_index(ix, a.W);
            }
            const int ch = kc * KC + k8 * 8;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (iy >= 0 && iy < a.H && ix >= 0 && ix < a.W && ch < a.Cin)
                v = *reinterpret_cast<const uint4*>(a.x + ((size_t)(n * a.H + iy) * a.W + ix) * a.x_pitch + a.x_coff + ch);
            *reinterpret_cast<uint4*>(As + (size_t)k8 * a.plane_bytes + pix * 16) = v;
        }
        // ---- B: packed weights of (nchunk, kc): one contiguous block
        {
            const uint4* src = reinterpret_cast<const uint4*>(a.wp + ((size_t)nchunk * a.p.kchunks + kc) * taps * KC * NC);
            uint4* dstp = reinterpret_cast<uint4*>(Bs);
            const int n16 = taps * KC * NC / 8;
            for (int idx = tid; idx < n16; idx += TC_THREADS) dstp[idx] = src[idx];
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t a_base = smem_u32(As), b_base = smem_u32(Bs);
            for (int tap = 0; tap < taps; ++tap) {
                const int r = tap / a.S, s = tap % a.S;
                for (int ks = 0; ks < ksteps; ++ks) {
                    const uint64_t ad = smem_desc(a_base + (2 * ks) * a.plane_bytes + (r * a.halo_w + s) * 16, a.plane_bytes, a.halo_w * 16);
                    const uint64_t bd = smem_desc(b_base + (tap * (KC / 8) + 2 * ks) * NC * 16, NC * 16, 128);
                    umma_bf16(tmem, ad, bd, idesc, (kc | tap | ks) != 0);
                }
            }
            umma_commit(smem_u32(&mbar));
        }
        mbar_wait(smem_u32(&mbar), phase);
        phase ^= 1;
    }
    tc_fence_after();

    // ---- epilogue: TMEM lane = output pixel, columns = output channels
    const int m = warp * 32 + lane;
    const int oy = y0 + m / TILE_W, ox = x0 + m % TILE_W;
    const bool pvalid = oy < a.OH && ox < a.OW;
    bf16* yp = a.y + ((size_t)(n * a.OH + (pvalid ? oy : 0)) * a.OW + (pvalid ? ox : 0)) * a.y_pitch + a.y_coff;
    const int c_begin = nchunk * NC;
    for (int g = 0; g < NC / 8; ++g) {
        const int c = c_begin + g * 8;
        float v[8];
        tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + g * 8, v);     // warp-collective: before any divergence
        const bool cvalid = c < a.Cout;
        if (a.bias && cvalid) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += a.bias[c + i];
        }
        if (a.stats) {
            float s1[8], s2[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { s1[i] = pvalid ? v[i] : 0.f; s2[i] = s1[i] * s1[i]; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    s1[i] += __shfl_xor_sync(0xffffffffu, s1[i], o);
                    s2[i] += __shfl_xor_sync(0xffffffffu, s2[i], o);
                }
            }
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) { part[warp * NC + g * 8 + i] = s1[i]; part[(4 + warp) * NC + g * 8 + i] = s2[i]; }
            }
        }
        if (pvalid && cvalid) {
            if (a.accumulate) {
                float o[8];
                Vec8<bf16>::load(yp + c, o);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] += o[i];
            }
            Vec8<bf16>::store(yp + c, v);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (a.stats) {
        for (int i = tid; i < NC; i += TC_THREADS) {
            const int c = c_begin + i;
            if (c < a.Cout) {
                const float t1 = ((part[i] + part[NC + i]) + part[2 * NC + i]) + part[3 * NC + i];
                const float t2 = ((part[4 * NC + i] + part[5 * NC + i]) + part[6 * NC + i]) + part[7 * NC + i];
                double* st = a.stats + (size_t)n * a.stats_nstride + c;
                atomicAdd(st, (double)t1);
                atomicAdd(st + a.stats_cstride, (double)t2);
            }
        }
    }
    if (warp == 0) tmem_dealloc<COLS>(tmem);
}

// ---- weight gradient on the tensor cores ------------------------------------------------------------------
// dW[tap][ci][co] = sum_pixels x[pixel+tap][ci] * dy[pixel][co]  as  D[M=co][N=ci] += A[co][k] * B[ci][k], k = pixels.
// Both operands are read straight from their NHWC channel-group planes as MN-MAJOR UMMA operands (the same
// shared-memory image the forward kernel uses as K-major): 8 consecutive pixels of a row are the 8 K-rows of a
// core matrix, LBO = next pixel row, SBO = next channel-group plane.  Each tap owns a block of TMEM columns.
//
// Persistent, warp-specialised: 4 producer warps stream (dy tile, x halo tile) pairs through a STAGES-deep
// cp.async ring; one elected thread of a 5th warp issues the tcgen05.mma's (K = 16 pixels each) and releases
// ring slots with tcgen05.commit; the accumulator stays in TMEM over ALL pixel tiles of the CTA (split-K over
// pixels) and is flushed once with coalesced fp32 reductions into the HWIO gradient.
constexpr int WG_PRODUCERS = 128;
constexpr int WG_THREADS = WG_PRODUCERS + 32;

struct WgPlan { int NCH, nchunks, tpp, tapgroups, mtiles, cols, splits, stage_bytes, planes_a_max, stages, smem_bytes; };

static inline WgPlan wg_plan(int Cin, int Cout, int taps, long long total_tiles, int plane_a, int plane_b) {
    WgPlan p;
    const int cin16 = (Cin + 15) / 16 * 16;
    p.NCH = cin16 < 128 ? cin16 : 128;          // keeps a stage (dy tile + x halo) below 80 KB
    p.nchunks = (cin16 + p.NCH - 1) / p.NCH;
    int tpp = 512 / p.NCH;
    if (tpp > taps) tpp = taps;
    p.tpp = tpp;
    p.tapgroups = (taps + p.tpp - 1) / p.tpp;
    p.mtiles = (Cout + 127) / 128;
    const int c = p.tpp * p.NCH;
    p.cols = c <= 32 ? 32 : (c <= 64 ? 64 : (c <= 128 ? 128 : (c <= 256 ? 256 : 512)));
    p.planes_a_max = Cout >= 128 ? 16 : (Cout + 7) / 8;
    p.stage_bytes = p.planes_a_max * plane_a + (p.NCH / 8) * plane_b;
    // the M=128 MMA always reads 16 A planes; planes beyond Cout are garbage rows that are never flushed, but the
    // reads must stay inside the allocation
    const int slack = 16 * plane_a > p.stage_bytes ? 16 * plane_a - p.stage_bytes : 0;
    p.stages = (size_t)p.stage_bytes * 3 + slack <= 200 * 1024 ? 3 : 2;
    p.smem_bytes = p.stage_bytes * p.stages + slack;
    const int work = p.mtiles * p.tapgroups * p.nchunks;
    const int per_sm = (p.cols <= 256 && p.smem_bytes <= 100 * 1024) ? 2 : 1;
    long long s = (148LL * per_sm) / work;
    if (s > total_tiles) s = total_tiles;
    if (s < 1) s = 1;
    p.splits = (int)s;
    return p;
}

struct WgArgs {
    int N, H, W, OH, OW, Cin, Cout, R, S, pad_t, pad_l, pad_mode;
    const bf16* x; int x_pitch, x_coff;
    const bf16* dy; int dy_pitch, dy_coff;
    float* dw;
    int tiles_x, tiles_y, total_tiles;
    WgPlan p;
    int plane_a, plane_b, halo_h, halo_w;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;      // src-size 0 -> the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <int COLS, int WG_STAGES>
__global__ void __launch_bounds__(WG_THREADS) wgrad_tc_kernel(const WgArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[WG_STAGES], empty_bar[WG_STAGES], done_bar;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int taps = a.R * a.S;
    const int NCH = a.p.NCH;
    int wi = blockIdx.y;
    const int nc = wi % a.p.nchunks; wi /= a.p.nchunks;
    const int tg = wi % a.p.tapgroups; wi /= a.p.tapgroups;
    const int mt = wi;
    const int tap0 = tg * a.p.tpp;
    const int ntaps = min(a.p.tpp, taps - tap0);
    const int co0 = mt * 128, ci0 = nc * NCH;
    const int planes_a = min(16, (a.Cout - co0 + 7) / 8);
    const int planes_b = min(NCH / 8, (a.Cin - ci0 + 7) / 8);
    const int halo_pix = a.halo_h * a.halo_w;
    const int b_off = a.p.planes_a_max * a.plane_a;                 // B planes follow the A planes inside a stage
    const int ntiles = (a.total_tiles - (int)blockIdx.x + a.p.splits - 1) / a.p.splits;   // tiles of this CTA

    if (warp == 4) tmem_alloc<COLS>(smem_u32(&tmem_slot));
    if (tid == 0) {
        for (int i = 0; i < WG_STAGES; ++i) { mbar_init(smem_u32(&full_bar[i]), WG_PRODUCERS); mbar_init(smem_u32(&empty_bar[i]), 1); }
        mbar_init(smem_u32(&done_bar), 1);
    }
    // N-side planes beyond Cin are never loaded: keep them finite (they only feed columns that are never flushed)
    for (int st = 0; st < WG_STAGES; ++st)
        for (int idx = tid; idx < (NCH / 8 - planes_b) * halo_pix; idx += WG_THREADS) {
            const int k8 = planes_b + idx / halo_pix, pix = idx % halo_pix;
            *reinterpret_cast<uint4*>(smem + (size_t)st * a.p.stage_bytes + b_off + (size_t)k8 * a.plane_b + pix * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t smem_base = smem_u32(smem);

    if (warp < 4) {
        // ===================== producers =====================
        for (int it = 0; it < ntiles; ++it) {
            const int stage = it % WG_STAGES;
            mbar_wait(smem_u32(&empty_bar[stage]), ((it / WG_STAGES) & 1) ^ 1);
            const int t = blockIdx.x + it * a.p.splits;
            const int n = t / (a.tiles_x * a.tiles_y);
            const int trem = t % (a.tiles_x * a.tiles_y);
            const int y0 = (trem / a.tiles_x) * TILE_H, x0 = (trem % a.tiles_x) * TILE_W;
            const uint32_t sbase = smem_base + stage * a.p.stage_bytes;
            {   // A: dy tile, one pixel per producer thread, all of its channel-group planes
                const int oy = y0 + tid / TILE_W, ox = x0 + tid % TILE_W;
                const bool v = oy < a.OH && ox < a.OW;
                const bf16* src = a.dy + ((size_t)(n * a.OH + (v ? oy : 0)) * a.OW + (v ? ox : 0)) * a.dy_pitch + a.dy_coff + co0;
                const uint32_t dst = sbase + tid * 16;
                for (int k8 = 0; k8 < planes_a; ++k8) cp_async16(dst + k8 * a.plane_a, src + k8 * 8, v);
            }
            for (int pix = tid; pix < halo_pix; pix += WG_PRODUCERS) {   // B: x halo tile
                const int hy = pix / a.halo_w, hx = pix % a.halo_w;
                int iy = y0 - a.pad_t + hy, ix = x0 - a.pad_l + hx;
                if (a.pad_mode == SEMB_PAD_REFLECT) {
                    if (iy > -a.H && iy < 2 * a.H - 1) iy = reflect_index(iy, a.H);
                    if (ix > -a.W && ix < 2 * a.W - 1) ix = reflect_index(ix, a.W);
                }
                const bool v = iy >= 0 && iy < a.H && ix >= 0 && ix < a.W;
                const bf16* src = a.x + ((size_t)(n * a.H + (v ? iy : 0)) * a.W + (v ? ix : 0)) * a.x_pitch + a.x_coff + ci0;
                const uint32_t dst = sbase + b_off + pix * 16;
                for (int k8 = 0; k8 < planes_b; ++k8) cp_async16(dst + k8 * a.plane_b, src + k8 * 8, v);
            }
            cp_async_commit();
            if (it >= WG_STAGES - 1) {          // the group issued WG_STAGES-1 iterations ago has landed
                cp_async_wait<WG_STAGES - 1>();
                fence_proxy_async();
                mbar_arrive(smem_u32(&full_bar[(it - (WG_STAGES - 1)) % WG_STAGES]));
            }
        }
        cp_async_wait<0>();
        fence_proxy_async();
        for (int it = max(0, ntiles - (WG_STAGES - 1)); it < ntiles; ++it) mbar_arrive(smem_u32(&full_bar[it % WG_STAGES]));
    } else if (lane == 0) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = instr_desc(128, NCH, 1, 1);
        const uint64_t ad0 = smem_desc(smem_base, TILE_W * 16, a.plane_a);
        const uint64_t bd0 = smem_desc(smem_base + b_off, a.halo_w * 16, a.plane_b);
        for (int it = 0; it < ntiles; ++it) {
            const int stage = it % WG_STAGES;
            mbar_wait(smem_u32(&full_bar[stage]), (it / WG_STAGES) & 1);
            tc_fence_after();
            const uint32_t soff = (uint32_t)(stage * a.p.stage_bytes) >> 4;
            for (int tl = 0; tl < ntaps; ++tl) {
                const int tap = tap0 + tl, r = tap / a.S, s = tap % a.S;
#pragma unroll
                for (int ks = 0; ks < TILE_H / 2; ++ks) {       // 16 pixels (two rows of 8) per MMA
                    const uint64_t ad = ad0 + soff + ks * (2 * TILE_W);
                    const uint64_t bd = bd0 + soff + ((2 * ks + r) * a.halo_w + s);
                    umma_bf16(tmem + tl * NCH, ad, bd, idesc, (it | ks) != 0);
                }
            }
            umma_commit(smem_u32(&empty_bar[stage]));
        }
        umma_commit(smem_u32(&done_bar));
    }
    // ---- flush: TMEM lane = co, columns = [tap][ci]; lanes of a warp hit consecutive co -> coalesced reductions
    if (warp < 4) {
        mbar_wait(smem_u32(&done_bar), 0);
        tc_fence_after();
        if (ntiles > 0) {
            const int co = co0 + warp * 32 + lane;
            for (int tl = 0; tl < ntaps; ++tl) {
                const int tap = tap0 + tl;
                for (int g = 0; g < NCH / 8; ++g) {
                    float v[8];
                    tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + tl * NCH + g * 8, v);
                    if (co < a.Cout) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int ci = ci0 + g * 8 + i;
                            if (ci < a.Cin) atomicAdd(a.dw + ((size_t)tap * a.Cin + ci) * a.Cout + co, v[i]);
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<COLS>(tmem);
}

static size_t tc_smem_bytes(const TcPlan& p, int taps, int plane_bytes) {
    return (size_t)(p.KC / 8) * plane_bytes + (size_t)taps * p.KC * p.NC * 2 + (size_t)8 * p.NC * sizeof(float);
}

}  // namespace semb

using namespace semb;

extern "C" int64_t semb_pack_weights_tc(const float* w, int32_t R, int32_t S, int32_t Cin, int32_t Cout, int32_t flip,
                                        void* dst, void* stream) {
    if (!((R == 1 && S == 1) || (R == 3 && S == 3)) || Cin <= 0 || Cout <= 0 || Cin % 8 || Cout % 8) {
        set_error("pack_weights_tc: need 1x1 or 3x3 and 8-padded channels (got %dx%d, %d->%d)", R, S, Cin, Cout);
        return SEMB_ESHAPE;
    }
    const int K = flip ? Cout : Cin, N = flip ? Cin : Cout;
    const TcPlan p = tc_plan(K, N, R * S);
    const long long total = (long long)p.nchunks * p.kchunks * R * S * p.KC * p.NC;
    if (!dst) return total * 2;
    if (!w) { set_error("pack_weights_tc: null weights"); return SEMB_ESHAPE; }
    long long blocks = cdivl(total, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    pack_weights_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(w, R, S, Cin, Cout, flip, p, reinterpret_cast<bf16*>(dst), total);
    const int rc = check_launch("pack_weights_tc");
    return rc ? rc : total * 2;
}

extern "C" int semb_conv2d_fwd_tc(const semb_conv_geom* g, const semb_tensor* x, const void* w_packed, const float* bias,
                                  const semb_tensor* y, void* stats, int32_t stats_nstride, int32_t stats_cstride,
                                  int32_t accumulate, void* stream) {
    SEMB_REQUIRE(g && x && y && w_packed, SEMB_ESHAPE, "conv_tc: null argument");
    SEMB_REQUIRE(g->dtype == SEMB_BF16, SEMB_ESHAPE, "conv_tc: bf16 storage only");
    SEMB_REQUIRE(g->stride == 1 && ((g->R == 1 && g->S == 1) || (g->R == 3 && g->S == 3)), SEMB_ESHAPE,
                 "conv_tc: stride-1 1x1 / 3x3 only (got %dx%d stride %d)", g->R, g->S, g->stride);
    SEMB_REQUIRE(view_ok(x) && view_ok(y) && x->C == g->Cin && y->C == g->Cout, SEMB_EALIGN, "conv_tc: bad tensor views");
    SEMB_REQUIRE(g->N > 0 && g->OH > 0 && g->OW > 0 && g->pad_t >= 0 && g->pad_l >= 0 && g->pad_t < g->R + TILE_H && g->pad_l < g->S + TILE_W,
                 SEMB_ESHAPE, "conv_tc: bad geometry");
    TcArgs a{};
    a.N = g->N; a.H = g->H; a.W = g->W; a.OH = g->OH; a.OW = g->OW; a.Cin = g->Cin; a.Cout = g->Cout;
    a.R = g->R; a.S = g->S; a.pad_t = g->pad_t; a.pad_l = g->pad_l; a.pad_mode = g->pad_mode;
    a.x = reinterpret_cast<const bf16*>(x->ptr); a.x_pitch = x->pitch; a.x_coff = x->coff;
    a.y = reinterpret_cast<bf16*>(y->ptr); a.y_pitch = y->pitch; a.y_coff = y->coff;
    a.wp = reinterpret_cast<const bf16*>(w_packed); a.bias = bias;
    a.stats = reinterpret_cast<double*>(stats); a.stats_nstride = stats_nstride; a.stats_cstride = stats_cstride;
    a.accumulate = accumulate;
    a.tiles_x = cdiv(g->OW, TILE_W); a.tiles_y = cdiv(g->OH, TILE_H);
    a.p = tc_plan(g->Cin, g->Cout, g->R * g->S);
    a.halo_h = TILE_H + g->R - 1; a.halo_w = TILE_W + g->S - 1;
    a.plane_bytes = a.halo_h * a.halo_w * 16 + 16;          // +16: consecutive planes start in different banks
    const size_t smem = tc_smem_bytes(a.p, g->R * g->S, a.plane_bytes);
    SEMB_REQUIRE(smem <= 200 * 1024, SEMB_EWORKSPACE, "conv_tc: %zu bytes of shared memory needed", smem);
    dim3 grid(g->N * a.tiles_x * a.tiles_y, a.p.nchunks);
    cudaError_t e = cudaSuccess;
#define SEMB_TC_LAUNCH(COLS)                                                                                         \
    e = cudaFuncSetAttribute(conv_tc_kernel<COLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
    if (e == cudaSuccess) conv_tc_kernel<COLS><<<grid, TC_THREADS, smem, as_stream(stream)>>>(a);
    switch (a.p.tmem_cols) {
        case 32: SEMB_TC_LAUNCH(32) break;
        case 64: SEMB_TC_LAUNCH(64) break;
        case 128: SEMB_TC_LAUNCH(128) break;
        default: SEMB_TC_LAUNCH(256) break;
    }
#undef SEMB_TC_LAUNCH
    if (e != cudaSuccess) { set_error("conv_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return SEMB_ECUDA; }
    return check_launch("conv_tc");
}

extern "C" int semb_conv2d_wgrad_tc(const semb_conv_geom* g, const semb_tensor* x, const semb_tensor* dy, float* dw,
                                    void* stream) {
    SEMB_REQUIRE(g && x && dy && dw, SEMB_ESHAPE, "wgrad_tc: null argument");
    SEMB_REQUIRE(g->dtype == SEMB_BF16, SEMB_ESHAPE, "wgrad_tc: bf16 storage only");
    SEMB_REQUIRE(g->stride == 1 && ((g->R == 1 && g->S == 1) || (g->R == 3 && g->S == 3)), SEMB_ESHAPE,
                 "wgrad_tc: stride-1 1x1 / 3x3 only (got %dx%d stride %d)", g->R, g->S, g->stride);
    SEMB_REQUIRE(view_ok(x) && view_ok(dy) && x->C == g->Cin && dy->C == g->Cout, SEMB_EALIGN, "wgrad_tc: bad tensor views");
    WgArgs a{};
    a.N = g->N; a.H = g->H; a.W = g->W; a.OH = g->OH; a.OW = g->OW; a.Cin = g->Cin; a.Cout = g->Cout;
    a.R = g->R; a.S = g->S; a.pad_t = g->pad_t; a.pad_l = g->pad_l; a.pad_mode = g->pad_mode;
    a.x = reinterpret_cast<const bf16*>(x->ptr); a.x_pitch = x->pitch; a.x_coff = x->coff;
    a.dy = reinterpret_cast<const bf16*>(dy->ptr); a.dy_pitch = dy->pitch; a.dy_coff = dy->coff;
    a.dw = dw;
    a.tiles_x = cdiv(g->OW, TILE_W); a.tiles_y = cdiv(g->OH, TILE_H);
    a.total_tiles = g->N * a.tiles_x * a.tiles_y;
    a.halo_h = TILE_H + g->R - 1; a.halo_w = TILE_W + g->S - 1;
    a.plane_a = TILE_H * TILE_W * 16 + 16;
    a.plane_b = a.halo_h * a.halo_w * 16 + 16;
    a.p = wg_plan(g->Cin, g->Cout, g->R * g->S, a.total_tiles, a.plane_a, a.plane_b);
    const size_t smem = (size_t)a.p.smem_bytes;
    SEMB_REQUIRE(smem <= 220 * 1024, SEMB_EWORKSPACE, "wgrad_tc: %zu bytes of shared memory needed", smem);
    dim3 grid(a.p.splits, a.p.mtiles * a.p.tapgroups * a.p.nchunks);
    cudaError_t e = cudaSuccess;
#define SEMB_WG_LAUNCH(COLS)                                                                                          \
    if (a.p.stages == 3) {                                                                                            \
        e = cudaFuncSetAttribute(wgrad_tc_kernel<COLS, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
        if (e == cudaSuccess) wgrad_tc_kernel<COLS, 3><<<grid, WG_THREADS, smem, as_stream(stream)>>>(a);             \
    } else {                                                                                                          \
        e = cudaFuncSetAttribute(wgrad_tc_kernel<COLS, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
        if (e == cudaSuccess) wgrad_tc_kernel<COLS, 2><<<grid, WG_THREADS, smem, as_stream(stream)>>>(a);             \
    }
    switch (a.p.cols) {
        case 32: SEMB_WG_LAUNCH(32) break;
        case 64: SEMB_WG_LAUNCH(64) break;
        case 128: SEMB_WG_LAUNCH(128) break;
        case 256: SEMB_WG_LAUNCH(256) break;
        default: SEMB_WG_LAUNCH(512) break;
    }
#undef SEMB_WG_LAUNCH
    if (e != cudaSuccess) { set_error("wgrad_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return SEMB_ECUDA; }
    return check_launch("wgrad_tc");
}
