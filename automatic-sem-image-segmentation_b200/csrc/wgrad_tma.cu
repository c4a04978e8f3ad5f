// Weight gradient of the zero-padded stride-1 3x3 layers on tcgen05 with TMA-staged operands.
//
//   dW[r][s][ci][co] += sum_pixels x[p + (r,s)][ci] * dy[p][co]
//
// Same formulation as wgrad_tc_stacked_kernel (conv_tc.cu): per 16x8 pixel tile the M side of the MMA is the x halo
// REPLICATED over the three horizontal taps -- plane (s, k8) = rows [y0-1, y0+17) x columns [x0-1+s, x0+7+s) of channel
// group k8, so one MN-major descriptor with a uniform group stride covers M = 3 * Cin_chunk <= 120 rows -- the vertical
// tap r is a 128-byte shift of the descriptor start, N = Cout_chunk, K = 16 pixels per MMA: 24 MMAs per tile.
// Differences (round-1 measurements: the cp.async version spent ~2.7 us per tile on address arithmetic for Cin = 32,
// the tensor pipe needs ~0.8 us):
//   * the replicas are written by the TMA unit: one 4-D box {8 channels, 8 pixels, 18 rows, 1} per (s, k8) plane, one
//     {8, 8, 16, 1} box per dy plane; zero padding and partial tiles are the tensor map's out-of-bounds fill;
//   * any Cin / Cout: input channels are processed in chunks of <= 40, output channels in chunks of <= 80 (grid.y);
//   * two MMA issuer warps split the K steps of a tile (even / odd), each with its own TMEM accumulators (summed in the
//     flush), because one converged warp issues an MMA only every ~130 cycles.
//
//   warps 0-3  flush (after the last tile): TMEM -> red.global.add.v4.f32 into the fp32 HWIO gradient
//   warps 4-5  MMA issuers
//   warp  6    TMA producer
//
// Replaces autograd's weight gradient of F.conv2d under keras.layers.Conv2D (UNet_Segmentation.py:421,465-468,490-499).
#include "tc_common.cuh"
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

namespace semb {

constexpr int WT_THREADS = 224;
constexpr int WT_MAX_STAGES = 6;
constexpr int WT_PLANE_A = (TILE_H + 2) * TILE_W * 16;   // 2304 bytes: [18 rows][8 pixels][8 channels]
constexpr int WT_PLANE_B = TILE_H * TILE_W * 16;         // 2048 bytes: [16 rows][8 pixels][8 channels]

struct WtArgs {
    int planar;              // operands are planar copies [N][C/8][H][W][8]: ONE TMA box per replica with 128-byte rows
    int ps;                  // planes between two horizontal-tap replicas in shared memory (p, or pc for planar boxes)
    int Cin, Cout, x_coff, dy_coff, pad_t, pad_l;
    float* dw;
    int tiles_x, tiles_y, total_tiles, splits;
    int pc, cin_chunks;      // planes (8 channels) per input-channel chunk, number of chunks
    int NB, cout_chunks;     // MMA N (multiple of 16) = output channels per chunk
    int a_bytes, stage_bytes, nstages, tmem_cols;
};

__device__ __forceinline__ void wt_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
                 "@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void wt_tma_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
        "@e cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n\t}"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}
__device__ __forceinline__ void wt_tmem_alloc(uint32_t slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void wt_tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

__global__ void __launch_bounds__(WT_THREADS) wgrad_tma_kernel(const WtArgs a, const __grid_constant__ CUtensorMap xmap,
                                                               const __grid_constant__ CUtensorMap dymap) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[WT_MAX_STAGES], empty_bar[WT_MAX_STAGES], done_bar;
    __shared__ uint32_t tmem_slot;
    pdl_trigger();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chunk = blockIdx.y;
    const int cic = chunk % a.cin_chunks, coc = chunk / a.cin_chunks;
    const int plane0 = cic * a.pc;                                   // first 8-channel plane of this CTA's input chunk
    const int p = min(a.pc, (a.Cin >> 3) - plane0);                  // planes in this chunk
    const int co0 = coc * a.NB;
    const int nb8 = min(a.NB, a.Cout - co0) >> 3;                    // real dy planes in this chunk
    const int NB = a.NB;
    const int ntiles = (a.total_tiles - (int)blockIdx.x + a.splits - 1) / a.splits;
    const uint32_t smem_base = (smem_u32(smem_raw) + 127u) & ~127u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));

    if (warp == 4) wt_tmem_alloc(smem_u32(&tmem_slot), (uint32_t)a.tmem_cols);
    if (tid == 0) {
        for (int i = 0; i < a.nstages; ++i) { mbar_init(smem_u32(&full_bar[i]), 1); mbar_init(smem_u32(&empty_bar[i]), 2); }
        mbar_init(smem_u32(&done_bar), 2);
    }
    {   // planes the TMA never writes (dy planes beyond Cout, rows read by the M = 128 MMA beyond 3p groups) must be finite
        const int n16 = (a.nstages * a.stage_bytes + 16 * WT_PLANE_A) >> 4;
        uint4* z = reinterpret_cast<uint4*>(smem);
        for (int i = tid; i < n16; i += WT_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    pdl_wait();

    if (warp == 6) {
        // ===================== TMA producer (whole warp converged) =====================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&dymap) : "memory");
        }
        __syncwarp();
        TileIter it(blockIdx.x, a.splits, a.tiles_x, a.tiles_y);
        int stage = 0;
        uint32_t phase = 1;
        // planar boxes always carry pc (resp. NB/8) planes; planes past the end of the tensor are zero-filled by the TMA
        const uint32_t bytes = a.planar ? (uint32_t)(3 * a.pc * WT_PLANE_A + (NB >> 3) * WT_PLANE_B)
                                        : (uint32_t)(3 * p * WT_PLANE_A + nb8 * WT_PLANE_B);
        for (int i = 0; i < ntiles; ++i, it.next()) {
            mbar_wait(smem_u32(&empty_bar[stage]), phase);
            const int y0 = it.ty * TILE_H, x0 = it.tx * TILE_W;
            const uint32_t sbase = smem_base + stage * a.stage_bytes;
            const uint32_t bar = smem_u32(&full_bar[stage]);
            wt_expect_tx(bar, bytes);
            if (a.planar) {
                // tensor map dims {W*8, H, C/8, N}: dim 0 runs over the 8-channel groups of one image row (16 B per pixel)
                wt_tma_4d(sbase + a.a_bytes, &dymap, x0 * 8, y0, co0 >> 3, it.n, bar);
                for (int s = 0; s < 3; ++s)
                    wt_tma_4d(sbase + s * a.ps * WT_PLANE_A, &xmap, (x0 - a.pad_l + s) * 8, y0 - a.pad_t, plane0, it.n, bar);
            } else {
                for (int k8 = 0; k8 < nb8; ++k8)
                    wt_tma_4d(sbase + a.a_bytes + k8 * WT_PLANE_B, &dymap, a.dy_coff + co0 + k8 * 8, x0, y0, it.n, bar);
                for (int s = 0; s < 3; ++s)
                    for (int k8 = 0; k8 < p; ++k8)
                        wt_tma_4d(sbase + (s * p + k8) * WT_PLANE_A, &xmap, a.x_coff + (plane0 + k8) * 8, x0 - a.pad_l + s, y0 - a.pad_t, it.n, bar);
            }
            if (++stage == a.nstages) { stage = 0; phase ^= 1; }
        }
    } else if (warp >= 4) {
        // ===================== MMA issuers: warp 4 even K steps, warp 5 odd K steps (converged, elect.sync) ==========
        const int mw = warp - 4;
        const uint32_t idesc = instr_desc(128, NB, 1, 1);
        const uint64_t ad0 = smem_desc(smem_base, TILE_W * 16, WT_PLANE_A);
        const uint64_t bd0 = smem_desc(smem_base + a.a_bytes, TILE_W * 16, WT_PLANE_B);
        const uint32_t a_hi = (uint32_t)(ad0 >> 32), b_hi = (uint32_t)(bd0 >> 32);
        const uint32_t a_lo0 = (uint32_t)ad0, b_lo0 = (uint32_t)bd0;
        const uint32_t stage16 = (uint32_t)a.stage_bytes >> 4;
        const uint32_t dbase = tmem + mw * 3 * NB;
        int stage = 0;
        uint32_t phase = 0;
        for (int i = 0; i < ntiles; ++i) {
            mbar_wait(smem_u32(&full_bar[stage]), phase);
            tc_fence_after();
            const uint32_t soff = (uint32_t)stage * stage16;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int kk = 0; kk < TILE_H / 4; ++kk) {           // this warp's K steps: ks = 2 * kk + mw
                    const uint32_t ks = 2 * kk + mw;
                    umma_bf16_elect(dbase + r * NB, a_lo0 + soff + (2 * ks + r) * TILE_W, a_hi,
                                    b_lo0 + soff + (2 * ks) * TILE_W, b_hi, idesc, (i | kk) != 0);
                }
            }
            umma_commit_elect(smem_u32(&empty_bar[stage]));
            if (++stage == a.nstages) { stage = 0; phase ^= 1; }
        }
        umma_commit_elect(smem_u32(&done_bar));
    }
    // ---- flush: TMEM lane = (s * p + k8) * 8 + c, columns = [issuer][r][co]; a thread owns one (s, ci) row of the gradient
    if (warp < 4) {
        mbar_wait_warp(smem_u32(&done_bar), 0);
        tc_fence_after();
        if (ntiles > 0) {
            const int l = warp * 32 + lane;
            const int g = l >> 3;
            const int ps = a.planar ? a.ps : p;
            const int s = g / ps, k8 = g - s * ps;
            const bool rvalid = s < 3 && k8 < p;
            const int ci = (plane0 + k8) * 8 + (l & 7);
            const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
            for (int r = 0; r < 3; ++r) {
                float* row = a.dw + ((size_t)(r * 3 + s) * a.Cin + ci) * a.Cout + co0;
                for (int j = 0; j < nb8; ++j) {
                    float v0[8], v1[8];
                    tmem_ld8(lane_addr + r * NB + j * 8, v0);
                    tmem_ld8(lane_addr + (3 + r) * NB + j * 8, v1);
                    if (rvalid) {       // 16-byte vector reductions: a quarter of the L2 atomic operations
                        red_add_v4(row + j * 8, v0[0] + v1[0], v0[1] + v1[1], v0[2] + v1[2], v0[3] + v1[3]);
                        red_add_v4(row + j * 8 + 4, v0[4] + v1[4], v0[5] + v1[5], v0[6] + v1[6], v0[7] + v1[7]);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) wt_tmem_dealloc(tmem, (uint32_t)a.tmem_cols);
}

// ---- host side --------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn wt_encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

static bool wt_make_map(EncodeTiledFn enc, CUtensorMap* map, const semb_tensor* t, int W, int H, int N, int box_h) {
    const cuuint64_t dims[4] = {(cuuint64_t)t->pitch, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)t->pitch * 2, (cuuint64_t)W * t->pitch * 2, (cuuint64_t)H * W * t->pitch * 2};
    const cuuint32_t box[4] = {8u, (cuuint32_t)TILE_W, (cuuint32_t)box_h, 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(t->ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// NHWC view (pitch, coff) -> planar [N][C/8][H][W][8]: the layout whose TMA boxes have 128-byte rows.  A block moves
// 32 pixels x 32 channel groups through shared memory so that both the reads (across channel groups of a pixel) and the
// writes (across pixels of a plane) are contiguous.
__global__ void __launch_bounds__(256) nhwc_to_planar_kernel(const bf16* __restrict__ src, int pitch, int coff, uint4* __restrict__ dst,
                                                             long long HW, int P) {
    __shared__ uint4 tile[32][33];
    pdl_trigger();
    pdl_wait();
    const int n = blockIdx.z;
    const long long px0 = (long long)blockIdx.x * 32;
    const int p0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;           // 32 x 8 threads
    for (int j = ty; j < 32; j += 8) {                                 // j = pixel, tx = plane
        const long long px = px0 + j;
        if (px < HW && p0 + tx < P)
            tile[j][tx] = *reinterpret_cast<const uint4*>(src + ((size_t)n * HW + px) * pitch + coff + (size_t)(p0 + tx) * 8);
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {                                 // j = plane, tx = pixel
        const long long px = px0 + tx;
        if (px < HW && p0 + j < P) dst[((size_t)n * P + p0 + j) * HW + px] = tile[tx][j];
    }
}

static bool wt_make_planar_map(EncodeTiledFn enc, CUtensorMap* map, const void* base, int W, int H, int P, int N, int box_h, int box_p) {
    const cuuint64_t dims[4] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)P, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)P * H * W * 16};
    const cuuint32_t box[4] = {(cuuint32_t)TILE_W * 8, (cuuint32_t)box_h, (cuuint32_t)box_p, 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

size_t wgrad_tma_workspace_bytes(const semb_conv_geom* g) {
    // planar copies of x and dy, each starting on a 128-byte boundary
    const size_t xb = ((size_t)g->N * g->H * g->W * g->Cin * 2 + 127) / 128 * 128;
    const size_t yb = ((size_t)g->N * g->OH * g->OW * g->Cout * 2 + 127) / 128 * 128;
    return xb + yb;
}

// Called by semb_conv2d_wgrad_tc (conv_tc.cu) after argument validation, for zero-padded 3x3 geometries.
// workspace != NULL: the operands are first copied into planar layout (see nhwc_to_planar_kernel) -- for the layers with
// hundreds of channels (CycleGAN residual blocks) the NHWC boxes (16-byte rows, one per pixel and channel group) made the
// TMA unit, not the tensor pipe, the limit of this kernel.
int wgrad_tma_launch(const semb_conv_geom* g, const semb_tensor* x, const semb_tensor* dy, float* dw, void* workspace, void* stream) {
    WtArgs a{};
    a.planar = workspace != nullptr;
    a.Cin = g->Cin; a.Cout = g->Cout; a.x_coff = x->coff; a.dy_coff = dy->coff; a.pad_t = g->pad_t; a.pad_l = g->pad_l;
    a.dw = dw;
    a.tiles_x = cdiv(g->OW, TILE_W); a.tiles_y = cdiv(g->OH, TILE_H);
    a.total_tiles = g->N * a.tiles_x * a.tiles_y;
    const int P = g->Cin / 8;
    a.cin_chunks = cdiv(P, 5);
    a.pc = cdiv(P, a.cin_chunks);
    a.cout_chunks = cdiv(g->Cout, 80);
    a.NB = (cdiv(g->Cout, a.cout_chunks) + 15) / 16 * 16;
    a.cout_chunks = cdiv(g->Cout, a.NB);
    a.ps = a.pc;
    a.a_bytes = 3 * a.pc * WT_PLANE_A;
    a.stage_bytes = a.a_bytes + (a.NB / 8) * WT_PLANE_B;
    const int need = 6 * a.NB;                                         // two issuers x three vertical taps
    a.tmem_cols = need <= 32 ? 32 : (need <= 64 ? 64 : (need <= 128 ? 128 : (need <= 256 ? 256 : 512)));
    const size_t slack = 16 * WT_PLANE_A + 128;                        // garbage-row reads of the last stage + alignment
    int nst = 2;
    while (nst < WT_MAX_STAGES && (size_t)(nst + 1) * a.stage_bytes + slack <= 104 * 1024) ++nst;
    static const int want_stages = [] { const char* e = getenv("SEMB_WGRAD_STAGES"); return e ? atoi(e) : 0; }();
    if (want_stages >= 2 && want_stages <= WT_MAX_STAGES) nst = want_stages;
    a.nstages = nst;
    const size_t smem = (size_t)nst * a.stage_bytes + slack;
    SEMB_REQUIRE(smem <= 220 * 1024, SEMB_EWORKSPACE, "wgrad_tma: %zu bytes of shared memory needed", smem);
    int per_sm = 512 / a.tmem_cols;
    if ((size_t)per_sm * (smem + 2048) > 220 * 1024) per_sm = (int)(220 * 1024 / (smem + 2048));
    if (per_sm > 2) per_sm = 2;
    if (per_sm < 1) per_sm = 1;
    const int chunks = a.cin_chunks * a.cout_chunks;
    // split-K over CTAs: every CTA ends with its share of the 9*Cin*Cout gradient reductions, so few-tile problems use
    // fewer CTAs: minimise tiles/splits * t_tile + splits * t_flush (t_tile ~ 1 us, ~60 G reduced floats per second)
    const double t_flush_us = 9.0 * (a.pc * 8) * a.NB / 60e3;
    long long splits = (long long)sqrt((double)a.total_tiles / t_flush_us);
    const long long cap = (148LL * per_sm + chunks - 1) / chunks;
    if (splits > cap) splits = cap;
    if (splits > a.total_tiles) splits = a.total_tiles;
    if (splits < 1) splits = 1;
    a.splits = (int)splits;

    EncodeTiledFn enc = wt_encode_tiled();
    SEMB_REQUIRE(enc != nullptr, SEMB_ECUDA, "wgrad_tma: cuTensorMapEncodeTiled is not available from the driver");
    CUtensorMap xmap, dymap;
    if (a.planar) {
        SEMB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) % 128) == 0, SEMB_EALIGN, "wgrad_tma: the workspace must be 128-byte aligned");
        uint8_t* xp = reinterpret_cast<uint8_t*>(workspace);
        uint8_t* yp = xp + ((size_t)g->N * g->H * g->W * g->Cin * 2 + 127) / 128 * 128;
        const long long hwx = (long long)g->H * g->W, hwy = (long long)g->OH * g->OW;
        const int Px = g->Cin / 8, Py = g->Cout / 8;
        launch_pdl(nhwc_to_planar_kernel, dim3((unsigned)cdivl(hwx, 32), cdiv(Px, 32), g->N), dim3(256), 0, as_stream(stream),
                   reinterpret_cast<const bf16*>(x->ptr), x->pitch, x->coff, reinterpret_cast<uint4*>(xp), hwx, Px);
        int rc = check_launch("nhwc_to_planar(x)");
        if (rc) return rc;
        launch_pdl(nhwc_to_planar_kernel, dim3((unsigned)cdivl(hwy, 32), cdiv(Py, 32), g->N), dim3(256), 0, as_stream(stream),
                   reinterpret_cast<const bf16*>(dy->ptr), dy->pitch, dy->coff, reinterpret_cast<uint4*>(yp), hwy, Py);
        rc = check_launch("nhwc_to_planar(dy)");
        if (rc) return rc;
        SEMB_REQUIRE(wt_make_planar_map(enc, &xmap, xp, g->W, g->H, Px, g->N, TILE_H + 2, a.pc) &&
                     wt_make_planar_map(enc, &dymap, yp, g->OW, g->OH, Py, g->N, TILE_H, a.NB / 8),
                     SEMB_ECUDA, "wgrad_tma: cuTensorMapEncodeTiled failed (planar)");
    } else {
        SEMB_REQUIRE(wt_make_map(enc, &xmap, x, g->W, g->H, g->N, TILE_H + 2) && wt_make_map(enc, &dymap, dy, g->OW, g->OH, g->N, TILE_H),
                     SEMB_ECUDA, "wgrad_tma: cuTensorMapEncodeTiled failed");
    }
    cudaError_t e = cudaFuncSetAttribute(wgrad_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("wgrad_tma: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return SEMB_ECUDA; }
    e = launch_pdl(wgrad_tma_kernel, dim3(a.splits, chunks), dim3(WT_THREADS), smem, as_stream(stream), a, xmap, dymap);
    if (e != cudaSuccess) { set_error("wgrad_tma: launch failed: %s", cudaGetErrorString(e)); cudaGetLastError(); return SEMB_ECUDA; }
    return check_launch("wgrad_tma");
}

}  // namespace semb
