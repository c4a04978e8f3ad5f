// Generic fp32-accumulate implicit-GEMM convolutions on the CUDA cores.
//
// This is the PARITY path (fp32 or bf16 storage, exact fp32 FMA accumulation) and the path for
// the layers that are not tensor-core work (Cin = 1 stems, Cout = 1 heads, strided / 4x4 / 7x7
// filters until their tcgen05 variants exist).  One kernel template covers
//   TR = false : y = conv(x, w)          F.conv2d            (UNet_Segmentation.py:421, CycleGAN.py:327..448)
//   TR = true  : dx = conv^T(dy, w)      its data gradient, and the forward of Conv2DTranspose
//                                        (UNet_Segmentation.py:542-551, CycleGAN.py:353)
// and a second one the weight gradient.  Tensors are NHWC with 8-padded channels (semb200.h).
#include "common.cuh"

namespace semb {

struct ConvArgs {
    int N, H, W, OH, OW, Cin, Cout, R, S, stride, pad_t, pad_l, pad_mode;
    const void* src; int src_pitch, src_coff;      // TR ? dy : x
    void* dst;       int dst_pitch, dst_coff;      // TR ? dx : y
    const float* w;
    const float* bias;
    double* stats; int stats_nstride, stats_cstride;
    int accumulate;
    int tiles_x, tiles_y;
};

constexpr int BK = 16;

template <typename T, int TH, int TW, int BN, bool TR>
__global__ void __launch_bounds__(256) conv_simt_kernel(const ConvArgs a) {
    constexpr int BM = TH * TW;
    constexpr int ELEMS = BM * BK / 256;          // A elements per thread per k-step (4, 8 or 16)
    constexpr int TPP = BK / ELEMS;               // threads per pixel in the A load
    constexpr int LDA = BM + 4;
    static_assert(BM * BN == 4096, "thread tile is 4x4 with 256 threads");
    __shared__ __align__(16) float As[BK][LDA];
    __shared__ __align__(16) float Bs[BK][BN];
    __shared__ float part[2048];                   // per-thread moment partials: [2][1024/BN][BN]

    const int tid = threadIdx.x;
    // output / source domains
    const int PH = TR ? a.H : a.OH, PW = TR ? a.W : a.OW;
    const int QH = TR ? a.OH : a.H, QW = TR ? a.OW : a.W;
    const int KC = TR ? a.Cout : a.Cin;           // reduction channels
    const int NC = TR ? a.Cin : a.Cout;           // produced channels

    const int tile = blockIdx.x;
    const int n = tile / (a.tiles_x * a.tiles_y);
    const int trem = tile % (a.tiles_x * a.tiles_y);
    const int py0 = (trem / a.tiles_x) * TH;
    const int px0 = (trem % a.tiles_x) * TW;
    const int nc0 = blockIdx.y * BN;

    // A-load role
    const int lp = tid / TPP;                     // pixel within tile
    const int lch = (tid % TPP) * ELEMS;          // first channel within k-step
    const int lpy = py0 + lp / TW, lpx = px0 + lp % TW;
    // compute role
    const int tx = tid % (BN / 4), ty = tid / (BN / 4);

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const T* src = reinterpret_cast<const T*>(a.src);

    for (int r = 0; r < a.R; ++r) {
        for (int s = 0; s < a.S; ++s) {
            // source pixel of this thread's A-load pixel for tap (r,s)
            int qy, qx;
            bool valid = (lpy < PH) && (lpx < PW);
            if (!TR) {
                qy = lpy * a.stride - a.pad_t + r;
                qx = lpx * a.stride - a.pad_l + s;
                if (a.pad_mode == SEMB_PAD_REFLECT) {
                    qy = reflect_index(qy, QH);
                    qx = reflect_index(qx, QW);
                }
                valid = valid && qy >= 0 && qy < QH && qx >= 0 && qx < QW;
            } else {
                const int ty_ = lpy + a.pad_t - r, tx_ = lpx + a.pad_l - s;
                valid = valid && ty_ >= 0 && tx_ >= 0 && (ty_ % a.stride) == 0 && (tx_ % a.stride) == 0;
                qy = ty_ / a.stride;
                qx = tx_ / a.stride;
                valid = valid && qy < QH && qx < QW;
            }
            const T* sp = src + ((size_t)(n * QH + (valid ? qy : 0)) * QW + (valid ? qx : 0)) * a.src_pitch + a.src_coff;
            const int tap = r * a.S + s;

            for (int c0 = 0; c0 < KC; c0 += BK) {
                // ---- stage A (BM pixels x 16 channels), transposed to k-major
#pragma unroll
                for (int e = 0; e < ELEMS; e += 4) {
                    float v[4] = {0.f, 0.f, 0.f, 0.f};
                    const int ch = c0 + lch + e;
                    if (valid && ch < KC) Vec4<T>::load(sp + ch, v);
#pragma unroll
                    for (int q = 0; q < 4; ++q) As[lch + e + q][lp] = v[q];
                }
                // ---- stage B (16 x BN weights)
                for (int i = tid; i < BK * BN; i += 256) {
                    int kk, nn;
                    if (!TR) { kk = i / BN; nn = i % BN; } else { nn = i / BK; kk = i % BK; }
                    const int kc = c0 + kk, nc = nc0 + nn;
                    float wv = 0.f;
                    if (kc < KC && nc < NC) {
                        wv = TR ? a.w[((size_t)tap * a.Cin + nc) * a.Cout + kc]
                                : a.w[((size_t)tap * a.Cin + kc) * a.Cout + nc];
                    }
                    Bs[kk][nn] = wv;
                }
                __syncthreads();
#pragma unroll
                for (int kk = 0; kk < BK; ++kk) {
                    const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
                    const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
                    const float aa[4] = {av.x, av.y, av.z, av.w};
                    const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
                }
                __syncthreads();
            }
        }
    }

    // ---- epilogue: bias, moments, store
    const int nc = nc0 + tx * 4;
    const bool ch_ok = nc < NC;
    float bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (a.bias && ch_ok) {
#pragma unroll
        for (int j = 0; j < 4; ++j) bv[j] = a.bias[nc + j];
    }
    float psum[4] = {0.f, 0.f, 0.f, 0.f}, psq[4] = {0.f, 0.f, 0.f, 0.f};
    T* dst = reinterpret_cast<T*>(a.dst);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = ty * 4 + i;
        const int oy = py0 + m / TW, ox = px0 + m % TW;
        if (oy < PH && ox < PW && ch_ok) {
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                v[j] = acc[i][j] + bv[j];
                psum[j] += v[j];
                psq[j] += v[j] * v[j];
            }
            T* dp = dst + ((size_t)(n * PH + oy) * PW + ox) * a.dst_pitch + a.dst_coff + nc;
            if (a.accumulate) {
                float o[4];
                Vec4<T>::load(dp, o);
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] += o[j];
            }
            Vec4<T>::store(dp, v);
        }
    }
    if (a.stats) {
        // Deterministic moments: per-thread partials go to shared memory, one thread per channel adds them in a
        // fixed order, and CTAs are combined with fp64 atomics (order-independent to ~1e-16).
        constexpr int NTY = 1024 / BN;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            part[ty * BN + tx * 4 + j] = psum[j];
            part[NTY * BN + ty * BN + tx * 4 + j] = psq[j];
        }
        __syncthreads();
        if (tid < BN && nc0 + tid < NC) {
            float s1 = 0.f, s2 = 0.f;
            for (int r2 = 0; r2 < NTY; ++r2) {
                s1 += part[r2 * BN + tid];
                s2 += part[NTY * BN + r2 * BN + tid];
            }
            double* st = a.stats + (size_t)n * a.stats_nstride + nc0 + tid;
            atomicAdd(st, (double)s1);
            atomicAdd(st + a.stats_cstride, (double)s2);
        }
    }
}

template <typename T, bool TR>
static int launch_conv(const ConvArgs& a0, cudaStream_t st) {
    ConvArgs a = a0;
    const int PH = TR ? a.H : a.OH, PW = TR ? a.W : a.OW;
    const int NC = TR ? a.Cin : a.Cout;
    // pick the tile by the number of produced channels
    if (NC <= 16) {
        a.tiles_x = cdiv(PW, 16); a.tiles_y = cdiv(PH, 16);
        dim3 grid(a.N * a.tiles_x * a.tiles_y, cdiv(NC, 16));
        conv_simt_kernel<T, 16, 16, 16, TR><<<grid, 256, 0, st>>>(a);
    } else if (NC <= 32) {
        a.tiles_x = cdiv(PW, 16); a.tiles_y = cdiv(PH, 8);
        dim3 grid(a.N * a.tiles_x * a.tiles_y, cdiv(NC, 32));
        conv_simt_kernel<T, 8, 16, 32, TR><<<grid, 256, 0, st>>>(a);
    } else {
        a.tiles_x = cdiv(PW, 8); a.tiles_y = cdiv(PH, 8);
        dim3 grid(a.N * a.tiles_x * a.tiles_y, cdiv(NC, 64));
        conv_simt_kernel<T, 8, 8, 64, TR><<<grid, 256, 0, st>>>(a);
    }
    return check_launch(TR ? "conv_simt<dgrad>" : "conv_simt<fwd>");
}

// ---------------------------------------------------------------------------------------------
// weight gradient: dw[tap][ci][co] += sum_p x[src(p,tap)][ci] * dy[p][co]
struct WgradArgs {
    int N, H, W, OH, OW, Cin, Cout, R, S, stride, pad_t, pad_l, pad_mode;
    const void* x; int x_pitch, x_coff;
    const void* dy; int dy_pitch, dy_coff;
    float* dw; float* dbias;
    int ppc;            // output pixels per CTA
    int co_tiles;
};

constexpr int WP = 32;  // pixels per smem stage

template <typename T, int BCI, int BCO>
__global__ void __launch_bounds__(256) wgrad_simt_kernel(const WgradArgs a) {
    constexpr int TG = (BCI / 4) * (BCO / 4);     // threads per pixel-split group
    constexpr int PS = 256 / TG;                  // pixel-split groups
    static_assert(TG * PS == 256, "tile");
    __shared__ __align__(16) float Xs[WP][BCI];
    __shared__ __align__(16) float Ds[WP][BCO];
    __shared__ float red[PS > 1 ? BCI * BCO : 1];

    const int tid = threadIdx.x;
    const int tap = blockIdx.z, r = tap / a.S, s = tap % a.S;
    const int ci0 = (blockIdx.y / a.co_tiles) * BCI;
    const int co0 = (blockIdx.y % a.co_tiles) * BCO;
    const long long P = (long long)a.N * a.OH * a.OW;
    const long long p_begin = (long long)blockIdx.x * a.ppc;
    const long long p_end = min(P, p_begin + a.ppc);

    const int tci = tid % (BCI / 4), tco = (tid / (BCI / 4)) % (BCO / 4), grp = tid / TG;
    float acc[4][4];
    float bsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const T* x = reinterpret_cast<const T*>(a.x);
    const T* dy = reinterpret_cast<const T*>(a.dy);
    const bool do_bias = a.dbias && tap == 0 && ci0 == 0;

    for (long long p0 = p_begin; p0 < p_end; p0 += WP) {
        // ---- stage x (WP pixels x BCI channels) and dy (WP x BCO)
        for (int i = tid; i < WP * (BCI / 4); i += 256) {
            const int pp = i / (BCI / 4), c4 = (i % (BCI / 4)) * 4;
            const long long p = p0 + pp;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (p < p_end && ci0 + c4 < a.Cin) {
                const int n = (int)(p / (a.OH * a.OW));
                const int rem = (int)(p % (a.OH * a.OW));
                int qy = (rem / a.OW) * a.stride - a.pad_t + r;
                int qx = (rem % a.OW) * a.stride - a.pad_l + s;
                if (a.pad_mode == SEMB_PAD_REFLECT) { qy = reflect_index(qy, a.H); qx = reflect_index(qx, a.W); }
                if (qy >= 0 && qy < a.H && qx >= 0 && qx < a.W)
                    Vec4<T>::load(x + ((size_t)(n * a.H + qy) * a.W + qx) * a.x_pitch + a.x_coff + ci0 + c4, v);
            }
            *reinterpret_cast<float4*>(&Xs[pp][c4]) = make_float4(v[0], v[1], v[2], v[3]);
        }
        for (int i = tid; i < WP * (BCO / 4); i += 256) {
            const int pp = i / (BCO / 4), c4 = (i % (BCO / 4)) * 4;
            const long long p = p0 + pp;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (p < p_end && co0 + c4 < a.Cout) Vec4<T>::load(dy + (size_t)p * a.dy_pitch + a.dy_coff + co0 + c4, v);
            *reinterpret_cast<float4*>(&Ds[pp][c4]) = make_float4(v[0], v[1], v[2], v[3]);
        }
        __syncthreads();
#pragma unroll 4
        for (int pp = grp; pp < WP; pp += PS) {
            const float4 av = *reinterpret_cast<const float4*>(&Xs[pp][tci * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Ds[pp][tco * 4]);
            const float aa[4] = {av.x, av.y, av.z, av.w};
            const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
            if (do_bias && tci == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) bsum[j] += bb[j];
            }
        }
        __syncthreads();
    }

    if (PS > 1) {
        for (int i = tid; i < BCI * BCO; i += 256) red[i] = 0.f;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd(&red[(tci * 4 + i) * BCO + tco * 4 + j], acc[i][j]);
        __syncthreads();
        for (int i = tid; i < BCI * BCO; i += 256) {
            const int ci = ci0 + i / BCO, co = co0 + i % BCO;
            if (ci < a.Cin && co < a.Cout) atomicAdd(a.dw + ((size_t)tap * a.Cin + ci) * a.Cout + co, red[i]);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ci = ci0 + tci * 4 + i, co = co0 + tco * 4 + j;
                if (ci < a.Cin && co < a.Cout) atomicAdd(a.dw + ((size_t)tap * a.Cin + ci) * a.Cout + co, acc[i][j]);
            }
    }
    if (do_bias && tci == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int co = co0 + tco * 4 + j;
            if (co < a.Cout) atomicAdd(a.dbias + co, bsum[j]);
        }
    }
}

template <typename T>
static int launch_wgrad(const WgradArgs& a0, cudaStream_t st) {
    WgradArgs a = a0;
    const long long P = (long long)a.N * a.OH * a.OW;
    a.ppc = 1024;
    const int chunks = (int)cdivl(P, a.ppc);
    const int taps = a.R * a.S;
    const int mx = a.Cin > a.Cout ? a.Cin : a.Cout;
    if (mx <= 16) {
        a.co_tiles = cdiv(a.Cout, 16);
        dim3 grid(chunks, cdiv(a.Cin, 16) * a.co_tiles, taps);
        wgrad_simt_kernel<T, 16, 16><<<grid, 256, 0, st>>>(a);
    } else if (mx <= 32) {
        a.co_tiles = cdiv(a.Cout, 32);
        dim3 grid(chunks, cdiv(a.Cin, 32) * a.co_tiles, taps);
        wgrad_simt_kernel<T, 32, 32><<<grid, 256, 0, st>>>(a);
    } else {
        a.co_tiles = cdiv(a.Cout, 64);
        dim3 grid(chunks, cdiv(a.Cin, 64) * a.co_tiles, taps);
        wgrad_simt_kernel<T, 64, 64><<<grid, 256, 0, st>>>(a);
    }
    return check_launch("wgrad_simt");
}

static int check_geom(const semb_conv_geom* g, const semb_tensor* in, const semb_tensor* out, bool tr) {
    SEMB_REQUIRE(g && in && out, SEMB_ESHAPE, "conv: null argument");
    SEMB_REQUIRE(g->N > 0 && g->H > 0 && g->W > 0 && g->OH > 0 && g->OW > 0, SEMB_ESHAPE, "conv: empty geometry");
    SEMB_REQUIRE(g->R > 0 && g->S > 0 && g->R <= 7 && g->S <= 7, SEMB_ESHAPE, "conv: kernel %dx%d unsupported", g->R, g->S);
    SEMB_REQUIRE(g->stride == 1 || g->stride == 2, SEMB_ESHAPE, "conv: stride %d unsupported", g->stride);
    SEMB_REQUIRE(g->dtype == SEMB_F32 || g->dtype == SEMB_BF16, SEMB_ESHAPE, "conv: bad dtype %d", g->dtype);
    SEMB_REQUIRE(view_ok(in) && view_ok(out), SEMB_EALIGN, "conv: tensor views must be 8-channel padded and 16B aligned");
    const semb_tensor* xin = tr ? out : in;
    const semb_tensor* yout = tr ? in : out;
    SEMB_REQUIRE(xin->C == g->Cin && yout->C == g->Cout, SEMB_ESHAPE, "conv: channel mismatch (Cin %d vs %d, Cout %d vs %d)",
                 g->Cin, xin->C, g->Cout, yout->C);
    // every output position must map inside the (padded) input
    const int need_h = (g->OH - 1) * g->stride - g->pad_t + g->R;
    const int need_w = (g->OW - 1) * g->stride - g->pad_l + g->S;
    if (g->pad_mode == SEMB_PAD_REFLECT) {
        SEMB_REQUIRE(!tr, SEMB_ESHAPE, "conv dgrad: reflect padding is folded by semb_pad_crop, not here");
        SEMB_REQUIRE(g->pad_t < g->H && g->pad_l < g->W && need_h - g->H < g->H && need_w - g->W < g->W, SEMB_ESHAPE,
                     "conv: reflect padding wider than the image");
    }
    return SEMB_OK;
}

}  // namespace semb

using namespace semb;

extern "C" int semb_conv2d_fwd(const semb_conv_geom* g, const semb_tensor* x, const float* w, const float* bias,
                               const semb_tensor* y, void* stats, int32_t stats_nstride, int32_t stats_cstride,
                               int32_t accumulate, void* stream) {
    int rc = check_geom(g, x, y, false);
    if (rc) return rc;
    SEMB_REQUIRE(w, SEMB_ESHAPE, "conv fwd: null weights");
    ConvArgs a{g->N, g->H, g->W, g->OH, g->OW, g->Cin, g->Cout, g->R, g->S, g->stride, g->pad_t, g->pad_l, g->pad_mode,
               x->ptr, x->pitch, x->coff, y->ptr, y->pitch, y->coff, w, bias, reinterpret_cast<double*>(stats), stats_nstride, stats_cstride,
               accumulate, 0, 0};
    return g->dtype == SEMB_BF16 ? launch_conv<bf16, false>(a, as_stream(stream)) : launch_conv<float, false>(a, as_stream(stream));
}

extern "C" int semb_conv2d_dgrad(const semb_conv_geom* g, const semb_tensor* dy, const float* w, const float* bias,
                                 const semb_tensor* dx, void* stats, int32_t stats_nstride, int32_t stats_cstride,
                                 int32_t accumulate, void* stream) {
    int rc = check_geom(g, dy, dx, true);
    if (rc) return rc;
    SEMB_REQUIRE(w, SEMB_ESHAPE, "conv dgrad: null weights");
    ConvArgs a{g->N, g->H, g->W, g->OH, g->OW, g->Cin, g->Cout, g->R, g->S, g->stride, g->pad_t, g->pad_l, SEMB_PAD_ZERO,
               dy->ptr, dy->pitch, dy->coff, dx->ptr, dx->pitch, dx->coff, w, bias, reinterpret_cast<double*>(stats), stats_nstride, stats_cstride,
               accumulate, 0, 0};
    return g->dtype == SEMB_BF16 ? launch_conv<bf16, true>(a, as_stream(stream)) : launch_conv<float, true>(a, as_stream(stream));
}

extern "C" int semb_conv2d_wgrad(const semb_conv_geom* g, const semb_tensor* x, const semb_tensor* dy, float* dw,
                                 float* dbias, void* stream) {
    int rc = check_geom(g, x, dy, false);
    if (rc) return rc;
    SEMB_REQUIRE(dw, SEMB_ESHAPE, "conv wgrad: null dw");
    WgradArgs a{g->N, g->H, g->W, g->OH, g->OW, g->Cin, g->Cout, g->R, g->S, g->stride, g->pad_t, g->pad_l, g->pad_mode,
                x->ptr, x->pitch, x->coff, dy->ptr, dy->pitch, dy->coff, dw, dbias, 0, 0};
    return g->dtype == SEMB_BF16 ? launch_wgrad<bf16>(a, as_stream(stream)) : launch_wgrad<float>(a, as_stream(stream));
}
