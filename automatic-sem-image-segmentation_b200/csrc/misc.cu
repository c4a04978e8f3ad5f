// Pooling, padding, losses, optimizer and boundary casts.  All HBM-bound, 8-channel vectors.
#include "common.cuh"

namespace semb {

struct PView { void* ptr; int pitch, coff; };
static inline PView pv(const semb_tensor* t) { return t ? PView{t->ptr, t->pitch, t->coff} : PView{nullptr, 0, 0}; }
template <typename T>
__device__ __forceinline__ T* at(const PView& v, size_t pixel, int c) {
    return reinterpret_cast<T*>(v.ptr) + pixel * v.pitch + v.coff + c;
}

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }

// ---- MaxPooling2D((2,2)) ------------------------------------------------------------------------
template <typename T, bool BWD>
__global__ void maxpool_kernel(PView x, PView y /* fwd: out, bwd: dy */, PView dx, int N, int H, int W, int C8, int acc) {
    pdl_trigger();
    pdl_wait();
    const int OH = H / 2, OW = W / 2;
    const long long total = (long long)N * OH * OW * C8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C8) * 8;
        const long long op = i / C8;
        const int ox = (int)(op % OW), oy = (int)((op / OW) % OH), n = (int)(op / ((long long)OW * OH));
        const size_t p00 = ((size_t)n * H + 2 * oy) * W + 2 * ox;
        float v[4][8];
        Vec8<T>::load(at<T>(x, p00, c), v[0]);
        Vec8<T>::load(at<T>(x, p00 + 1, c), v[1]);
        Vec8<T>::load(at<T>(x, p00 + W, c), v[2]);
        Vec8<T>::load(at<T>(x, p00 + W + 1, c), v[3]);
        if (!BWD) {
            float m[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) m[k] = fmaxf(fmaxf(v[0][k], v[1][k]), fmaxf(v[2][k], v[3][k]));
            Vec8<T>::store(at<T>(y, (size_t)op, c), m);
        } else {
            float g[8], d[4][8];
            Vec8<T>::load(at<T>(y, (size_t)op, c), g);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                int arg = 0;
                float best = v[0][k];
#pragma unroll
                for (int q = 1; q < 4; ++q)
                    if (v[q][k] > best) { best = v[q][k]; arg = q; }   // first maximum wins (ATen)
#pragma unroll
                for (int q = 0; q < 4; ++q) d[q][k] = (q == arg) ? g[k] : 0.f;
            }
            const size_t pos[4] = {p00, p00 + 1, p00 + W, p00 + W + 1};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                T* o = at<T>(dx, pos[q], c);
                if (acc) {
                    float old[8];
                    Vec8<T>::load(o, old);
#pragma unroll
                    for (int k = 0; k < 8; ++k) d[q][k] += old[k];
                }
                Vec8<T>::store(o, d[q]);
            }
        }
    }
}

// ---- reflect pad / crop / zero pad / reflect fold -------------------------------------------------
template <typename T>
__global__ void pad_crop_kernel(PView x, PView y, int N, int H, int W, int OH, int OW, int top, int left, int mode,
                                int C8, int acc) {
    pdl_trigger();
    pdl_wait();
    const long long total = (long long)N * OH * OW * C8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C8) * 8;
        const long long op = i / C8;
        const int ox = (int)(op % OW), oy = (int)((op / OW) % OH), n = (int)(op / ((long long)OW * OH));
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = 0.f;
        if (mode == 0) {
            const int sy = reflect_index(oy - top, H), sx = reflect_index(ox - left, W);
            Vec8<T>::load(at<T>(x, ((size_t)n * H + sy) * W + sx, c), v);
        } else if (mode == 1) {
            Vec8<T>::load(at<T>(x, ((size_t)n * H + oy + top) * W + ox + left, c), v);
        } else if (mode == 2) {
            const int sy = oy - top, sx = ox - left;
            if (sy >= 0 && sy < H && sx >= 0 && sx < W) Vec8<T>::load(at<T>(x, ((size_t)n * H + sy) * W + sx, c), v);
        } else {
            // gradient of reflect_pad: x is the padded-domain gradient (H,W); y the unpadded one (OH,OW)
            const int bot = H - OH - top, right = W - OW - left;
            int ys[3], xs[3], ny = 0, nx = 0;
            ys[ny++] = oy + top;
            if (oy >= 1 && oy <= top) ys[ny++] = top - oy;
            if (oy <= OH - 2 && oy >= OH - 1 - bot) ys[ny++] = top + 2 * (OH - 1) - oy;
            xs[nx++] = ox + left;
            if (ox >= 1 && ox <= left) xs[nx++] = left - ox;
            if (ox <= OW - 2 && ox >= OW - 1 - right) xs[nx++] = left + 2 * (OW - 1) - ox;
            for (int a = 0; a < ny; ++a)
                for (int b = 0; b < nx; ++b) {
                    float t[8];
                    Vec8<T>::load(at<T>(x, ((size_t)n * H + ys[a]) * W + xs[b], c), t);
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[k] += t[k];
                }
        }
        T* o = at<T>(y, (size_t)op, c);
        if (acc) {
            float old[8];
            Vec8<T>::load(o, old);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] += old[k];
        }
        Vec8<T>::store(o, v);
    }
}

// ---- losses -------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int K>
__device__ __forceinline__ void block_flush(float (&v)[K], float* out) {
    __shared__ float red[K][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        v[k] = warp_sum(v[k]);
        if (lane == 0) red[k][warp] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < K) {
        float s = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[threadIdx.x][w];
        atomicAdd(out + threadIdx.x, s);
    }
}

// weighted BCE on channel 0 of p (the head has one logical channel)
template <typename T>
__global__ void __launch_bounds__(256) loss_wbce_kernel(PView p, const float* __restrict__ yt, PView dp, long long count,
                                                        float weighting, float* out) {
    float acc[3] = {0.f, 0.f, 0.f};
    const float eps = 1e-7f, inv_count = 1.f / (float)count;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        float v[8];
        Vec8<T>::load(at<T>(p, (size_t)i, 0), v);
        const float pr = v[0], y = yt[i];
        const float pc = fminf(fmaxf(pr, eps), 1.f - eps);
        const float w = y * (weighting - 1.f) + 1.f;
        acc[0] += -w * (y * logf(pc) + (1.f - y) * logf(1.f - pc));
        acc[1] += fabsf(y - pr);
        acc[2] += ((pr > 0.5f ? 1.f : 0.f) == y) ? 1.f : 0.f;
        if (dp.ptr) {
            float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (pr >= eps && pr <= 1.f - eps) g[0] = w * (-(y / pc) + (1.f - y) / (1.f - pc)) * inv_count;
            Vec8<T>::store(at<T>(dp, (size_t)i, 0), g);
        }
    }
    block_flush<3>(acc, out);
}

// Same loss, but the probability is recomputed in fp32 from the head's PRE-activation u = z*scale + shift (z = the stored
// conv output, scale/shift = the head BatchNorm's affine).  In bf16 storage the stored sigmoid output has a spacing of 2^-8
// just below 1, so p > 0.998 would round to exactly 1: the Keras clip would then zero the gradient of a confidently wrong
// pixel and dL/dp = 1/(1-p) would be off by tens of percent near saturation.  Keras (fp32) only saturates past |u| ~ 16.
template <typename T>
__global__ void __launch_bounds__(256) loss_wbce_logits_kernel(PView z, const float* __restrict__ scale, const float* __restrict__ shift,
                                                               const float* __restrict__ yt, PView dp, long long count, float weighting,
                                                               float* out) {
    pdl_trigger();
    pdl_wait();
    float acc[3] = {0.f, 0.f, 0.f};
    const float eps = 1e-7f, inv_count = 1.f / (float)count;
    const float sc = scale[0], sh = shift[0];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
        float v[8];
        Vec8<T>::load(at<T>(z, (size_t)i, 0), v);
        const float u = fmaf(v[0], sc, sh);
        const float pr = 1.f / (1.f + expf(-u));
        const float y = yt[i];
        const float pc = fminf(fmaxf(pr, eps), 1.f - eps);
        const float w = y * (weighting - 1.f) + 1.f;
        acc[0] += -w * (y * logf(pc) + (1.f - y) * logf(1.f - pc));
        acc[1] += fabsf(y - pr);
        acc[2] += ((pr > 0.5f ? 1.f : 0.f) == y) ? 1.f : 0.f;
        if (dp.ptr) {
            float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (pr >= eps && pr <= 1.f - eps) g[0] = w * (-(y / pc) + (1.f - y) / (1.f - pc)) * inv_count;
            Vec8<T>::store(at<T>(dp, (size_t)i, 0), g);
        }
    }
    block_flush<3>(acc, out);
}

// L1 / L2 between a and b (or a constant target) over all 8-padded channels (pads are zero in both)
template <typename T>
__global__ void __launch_bounds__(256) loss_l1_l2_kernel(PView a, PView b, float target, int kind, long long n_pixels, int C8,
                                                         int c_logical, float gscale, PView da, int acc, float* out) {
    pdl_trigger();
    pdl_wait();
    float sum[1] = {0.f};
    const long long total = n_pixels * C8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C8) * 8;
        const size_t px = (size_t)(i / C8);
        float va[8], vb[8], g[8];
        Vec8<T>::load(at<T>(a, px, c), va);
        if (b.ptr) Vec8<T>::load(at<T>(b, px, c), vb);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const bool real = (c + k) < c_logical;
            const float d = real ? va[k] - (b.ptr ? vb[k] : target) : 0.f;
            if (kind == 0) { sum[0] += fabsf(d); g[k] = d > 0.f ? gscale : (d < 0.f ? -gscale : 0.f); }
            else { sum[0] += d * d; g[k] = 2.f * d * gscale; }
        }
        if (da.ptr) {
            T* o = at<T>(da, px, c);
            if (acc) {
                float old[8];
                Vec8<T>::load(o, old);
#pragma unroll
                for (int k = 0; k < 8; ++k) g[k] += old[k];
            }
            Vec8<T>::store(o, g);
        }
    }
    block_flush<1>(sum, out);
}

// ---- Adam ---------------------------------------------------------------------------------------------------
struct AdamState { long long t; float alpha; float pad; };

__global__ void adam_tick_kernel(AdamState* st, const float* lr, float b1, float b2) {
    pdl_trigger();
    pdl_wait();
    const long long t = st->t + 1;
    st->t = t;
    st->alpha = (float)((double)lr[0] * sqrt(1.0 - pow((double)b2, (double)t)) / (1.0 - pow((double)b1, (double)t)));
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long long n, const AdamState* st, float b1, float b2,
                                                   float eps, float gscale) {
    pdl_trigger();
    pdl_wait();
    const float alpha = st->alpha;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += (long long)gridDim.x * blockDim.x * 4) {
        if (i + 4 <= n) {
            float4 g4 = *reinterpret_cast<const float4*>(g + i);
            float4 m4 = *reinterpret_cast<float4*>(m + i), v4 = *reinterpret_cast<float4*>(v + i), w4 = *reinterpret_cast<float4*>(w + i);
            float gg[4] = {g4.x * gscale, g4.y * gscale, g4.z * gscale, g4.w * gscale};
            float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w}, ww[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                mm[k] += (gg[k] - mm[k]) * (1.f - b1);
                vv[k] += (gg[k] * gg[k] - vv[k]) * (1.f - b2);
                ww[k] -= alpha * mm[k] / (sqrtf(vv[k]) + eps);
            }
            *reinterpret_cast<float4*>(m + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
            *reinterpret_cast<float4*>(v + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
            *reinterpret_cast<float4*>(w + i) = make_float4(ww[0], ww[1], ww[2], ww[3]);
        } else {
            for (long long j = i; j < n; ++j) {
                const float gg = g[j] * gscale;
                m[j] += (gg - m[j]) * (1.f - b1);
                v[j] += (gg * gg - v[j]) * (1.f - b2);
                w[j] -= alpha * m[j] / (sqrtf(v[j]) + eps);
            }
        }
    }
}

__global__ void fill_kernel(float* p, long long n, float value) {
    pdl_trigger();
    pdl_wait();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = value;
}

template <typename T>
__global__ void cast_in_kernel(const float* __restrict__ src, int src_C, PView dst, long long n_pixels, int C8) {
    pdl_trigger();
    pdl_wait();
    const long long total = n_pixels * C8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C8) * 8;
        const size_t px = (size_t)(i / C8);
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = (c + k < src_C) ? src[px * src_C + c + k] : 0.f;
        Vec8<T>::store(at<T>(dst, px, c), v);
    }
}

template <typename T>
__global__ void cast_out_kernel(PView src, float* __restrict__ dst, int dst_C, long long n_pixels, int C8) {
    const long long total = n_pixels * C8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C8) * 8;
        const size_t px = (size_t)(i / C8);
        float v[8];
        Vec8<T>::load(at<T>(src, px, c), v);
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (c + k < dst_C) dst[px * dst_C + c + k] = v[k];
    }
}


// ---- depth-to-space for Conv2DTranspose(2x2, stride 2) expressed as a 1x1 conv with 4*C outputs ------------------
// dir 0: dst[n,2y+r,2x+s,c] = src[n,y,x,(2r+s)*C+c] + bias[c]        dir 1: src[n,y,x,(2r+s)*C+c] = dst[n,2y+r,2x+s,c]
template <typename T>
__global__ void pixel_shuffle2_kernel(PView src, PView dst, int N, int H, int W, int DH, int DW, int C8, const float* __restrict__ bias,
                                      int dir, int acc) {
    pdl_trigger();
    pdl_wait();
    // dst is (N, DH, DW, C) with DH <= 2H, DW <= 2W (odd sizes: the missing last row / column reads as zero, is not written)
    const long long total = (long long)N * H * W * 4 * C8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C8) * 8;
        long long t = i / C8;
        const int q = (int)(t % 4); t /= 4;
        const int x = (int)(t % W), y = (int)((t / W) % H), n = (int)(t / ((long long)W * H));
        const size_t sp = ((size_t)n * H + y) * W + x;
        const int dy = 2 * y + (q >> 1), dx = 2 * x + (q & 1);
        const bool inside = dy < DH && dx < DW;
        const size_t dp = ((size_t)n * DH + (inside ? dy : 0)) * DW + (inside ? dx : 0);
        float v[8];
        if (dir == 0) {
            if (!inside) continue;
            Vec8<T>::load(at<T>(src, sp, q * C8 * 8 + c), v);
            if (bias) {
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] += bias[c + k];
            }
            if (acc) {
                float o[8];
                Vec8<T>::load(at<T>(dst, dp, c), o);
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] += o[k];
            }
            Vec8<T>::store(at<T>(dst, dp, c), v);
        } else {
            if (inside) Vec8<T>::load(at<T>(dst, dp, c), v);
            else {
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = 0.f;
            }
            Vec8<T>::store(at<T>(src, sp, q * C8 * 8 + c), v);
        }
    }
}

// ---- UpSampling2D(size=(2,2)), nearest neighbour (CycleGAN.py:349, use_resize_convolution) --------------------------------
// dir 0: big[n,2y+r,2x+s,c] = small[n,y,x,c];   dir 1 (its gradient): small[n,y,x,c] (+)= sum_{r,s} big[n,2y+r,2x+s,c]
template <typename T>
__global__ void upsample2x_kernel(PView small, PView big, int N, int H, int W, int C8, int dir, int acc) {
    const long long total = (long long)N * H * W * C8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C8) * 8;
        const long long t = i / C8;
        const int x = (int)(t % W), y = (int)((t / W) % H), n = (int)(t / ((long long)W * H));
        const size_t sp = ((size_t)n * H + y) * W + x;
        const size_t bp = ((size_t)n * 2 * H + 2 * y) * (2 * W) + 2 * x;
        float v[8];
        if (dir == 0) {
            Vec8<T>::load(at<T>(small, sp, c), v);
            Vec8<T>::store(at<T>(big, bp, c), v);
            Vec8<T>::store(at<T>(big, bp + 1, c), v);
            Vec8<T>::store(at<T>(big, bp + 2 * W, c), v);
            Vec8<T>::store(at<T>(big, bp + 2 * W + 1, c), v);
        } else {
            float a[8], b[8], d[8], e[8];
            Vec8<T>::load(at<T>(big, bp, c), a);
            Vec8<T>::load(at<T>(big, bp + 1, c), b);
            Vec8<T>::load(at<T>(big, bp + 2 * W, c), d);
            Vec8<T>::load(at<T>(big, bp + 2 * W + 1, c), e);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = (a[k] + b[k]) + (d[k] + e[k]);
            if (acc) {
                float o[8];
                Vec8<T>::load(at<T>(small, sp, c), o);
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] += o[k];
            }
            Vec8<T>::store(at<T>(small, sp, c), v);
        }
    }
}

// ---- LeakyReLU x Dropout as one multiplicative mask (WGAN-GP critic, WassersteinGAN.py:547-621) -----------------------------
// y (+)= x * slope(z) * m with slope(z) = 1 for z > 0 else `neg` (z == NULL: 1) and m the dropout keep mask already scaled by
// 1/(1-rate) (m == NULL: 1).  With x = z it is LeakyReLU(z) * m (forward); with x = dy it is the gradient; with x = u it is
// the critic LINEARISED at z (the gradient-penalty tower: dP/dW = backprop of <grad_x D, u> through the same masks).
template <typename T>
__global__ void mask_mul_kernel(PView x, PView z, PView m, PView y, long long npix, int C8, float neg, int has_z, int has_m, int acc) {
    pdl_trigger();
    pdl_wait();
    const long long total = npix * C8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C8) * 8;
        const size_t p = (size_t)(i / C8);
        float v[8], t[8];
        Vec8<T>::load(at<T>(x, p, c), v);
        if (has_z) {
            Vec8<T>::load(at<T>(z, p, c), t);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] *= t[k] > 0.f ? 1.f : neg;
        }
        if (has_m) {
            Vec8<T>::load(at<T>(m, p, c), t);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] *= t[k];
        }
        if (acc) {
            Vec8<T>::load(at<T>(y, p, c), t);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] += t[k];
        }
        Vec8<T>::store(at<T>(y, p, c), v);
    }
}

// ---- gradient penalty of WGAN-GP (WassersteinGAN.py:88-121): one block per sample ------------------------------------------
// g = d(critic)/d(x_hat) (N, HW pixels, C channels);  norm_n = sqrt(sum g_n^2);  sums[0] += (norm_n - 1)^2, sums[1] += norm_n;
// u_n = scale * (norm_n - 1) / norm_n * g_n  = d(scale/2 * sum_n (norm_n - 1)^2) / d g_n   (scale = 2 * gp_weight / N).
template <typename T>
__global__ void __launch_bounds__(256) gp_direction_kernel(PView g, PView u, long long HW, int C8, float scale, float* __restrict__ sums) {
    __shared__ float red[8];
    __shared__ float s_coef;
    const int n = blockIdx.x;
    const long long total = HW * C8;
    float acc = 0.f;
    for (long long i = threadIdx.x; i < total; i += blockDim.x) {
        float v[8];
        Vec8<T>::load(at<T>(g, (size_t)(n * HW + i / C8), (int)(i % C8) * 8), v);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc = fmaf(v[k], v[k], acc);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        const float norm = sqrtf(t);
        atomicAdd(sums, (norm - 1.f) * (norm - 1.f));
        atomicAdd(sums + 1, norm);
        s_coef = norm > 0.f ? scale * (norm - 1.f) / norm : 0.f;
    }
    __syncthreads();
    const float coef = s_coef;
    for (long long i = threadIdx.x; i < total; i += blockDim.x) {
        float v[8];
        const size_t p = (size_t)(n * HW + i / C8);
        const int c = (int)(i % C8) * 8;
        Vec8<T>::load(at<T>(g, p, c), v);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] *= coef;
        Vec8<T>::store(at<T>(u, p, c), v);
    }
}

// Stride-2 convolutions on the stride-1 tensor-core kernels: a k x k (k = 3, 4; 5 with leading pad 1: a full 3x3) stride-2 conv over X equals a 2x2 stride-1
// conv over the space-to-depth image X' (H/2, W/2, 4C), embedded here in a 3x3 kernel (taps with r2 = 2 or s2 = 2 are zero):
//   w3[r2][s2][(2dy+dx)*Cin + ci][co] = w[2 r2 + dy - pt][2 s2 + dx - pl][ci][co]   (0 outside the k x k kernel)
// with pt, pl in {0, 1} the leading padding.  dir 0 writes w3 from the master weights, dir 1 ADDS the gradient of w3
// (written by the weight-gradient kernel) into the master gradient.
__global__ void s2d_weights_kernel(float* __restrict__ w, int k, int pt, int pl, int Cin, int Cout, float* __restrict__ w3, int dir) {
    const long long total = 9LL * 4 * Cin * Cout;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(i % Cout);
        long long t = i / Cout;
        const int kk = (int)(t % (4 * Cin));
        const int tap = (int)(t / (4 * Cin));
        const int q = kk / Cin, ci = kk - q * Cin;
        const int r2 = tap / 3, s2 = tap - 3 * r2;
        const int r = 2 * r2 + (q >> 1) - pt, s_ = 2 * s2 + (q & 1) - pl;
        // k = 3, 4 fill a 2x2 corner of the 3x3 kernel (r < k cuts the rest); k = 5 with a leading pad of 1 (Keras 'same' on even
        // sizes: input rows 2y-1 .. 2y+3 = row pairs y-1, y, y+1) fills all nine positions
        const bool valid = r >= 0 && r < k && s_ >= 0 && s_ < k;
        if (dir == 0) w3[i] = valid ? w[((size_t)(r * k + s_) * Cin + ci) * Cout + co] : 0.f;
        else if (valid) w[((size_t)(r * k + s_) * Cin + ci) * Cout + co] += w3[i];      // the map valid (tap, kk) -> (r, s, ci) is injective
    }
}

// res_path unit (UNet_Segmentation.py:490-499): the 3x3 conv and the 1x1 shortcut conv read the same tensor, so they run as
// ONE 3x3 conv with Ca + Cs outputs: w3[r][s][ci][0:Ca] = wa[r][s][ci][:], w3[1][1][ci][Ca:] = ws[ci][:], zero on the other
// taps of the shortcut columns (N = 16 -> 32 costs the tensor pipe nothing at these widths; one launch, one read of the
// input, one data gradient without a read-modify-write of dx).  dir 0 writes w3 from the masters, dir 1 ADDS the gradient
// of w3 into the two master gradients (the off-centre shortcut taps of dw3 are dropped: those weights do not exist).
__global__ void merge_weights_kernel(float* __restrict__ wa, float* __restrict__ ws, int Cin, int Ca, int Cs, float* __restrict__ w3, int dir) {
    const int Ct = Ca + Cs;
    const long long total = 9LL * Cin * Ct;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(i % Ct);
        const long long t = i / Ct;
        const int ci = (int)(t % Cin), tap = (int)(t / Cin);
        if (co < Ca) {
            float* m = wa + ((size_t)tap * Cin + ci) * Ca + co;
            if (dir == 0) w3[i] = *m; else *m += w3[i];
        } else {
            float* m = ws + (size_t)ci * Cs + (co - Ca);
            if (dir == 0) w3[i] = tap == 4 ? *m : 0.f;
            else if (tap == 4) *m += w3[i];
        }
    }
}

// stats[g][k][c] += sum_q temp[g][k][q*C + c]: moments of a depth-to-space output from the moments of its 4C-channel source
__global__ void fold_stats4_kernel(const double* __restrict__ temp, double* __restrict__ stats, int groups, int C, int stats_nstride,
                                   int stats_cstride) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= groups * 2 * C) return;
    const int c = i % C, k = (i / C) % 2, g = i / (2 * C);
    const double* t = temp + ((size_t)g * 2 + k) * 4 * C;
    stats[(size_t)g * stats_nstride + (size_t)k * stats_cstride + c] += ((t[c] + t[C + c]) + t[2 * C + c]) + t[3 * C + c];
}

// ---- 7x7 single-channel convs (CycleGAN generator stem 1->F and head F->1) as 1x1 tensor-core convs ------------------
// "big" is the reflect-padded domain (H+k-1, W+k-1), "small" the output domain (H, W); T8 = pad8(k*k) tap channels.
//   mode 0  patches[p][t]   = big[p + (r,s)][0]                       (stem forward: im2col of ONE channel)
//   mode 1  dbig[u][0]      = sum_t dpatches[u - (r,s)][t]            (stem data gradient, padded domain)
//   mode 2  small[p][0]     = bias + sum_t zbig[p + (r,s)][t]         (head forward: shift-and-add of the per-tap 1x1 conv)
//   mode 3  dzbig[u][t]     = dsmall[u - (r,s)][0]                    (head backward)
// channel lanes beyond the real ones are written as zeros.
template <typename T>
__global__ void tap_patch_kernel(PView small, PView big, int N, int H, int W, int k, int T8, const float* __restrict__ bias, int mode) {
    const int BH = H + k - 1, BW = W + k - 1, taps = k * k;
    if (mode == 0 || mode == 3) {
        // one thread = 8 tap channels of one pixel (mode 0: output pixel p of `small`-sized patches; mode 3: padded pixel u)
        const int OHh = mode == 0 ? H : BH, OWw = mode == 0 ? W : BW;
        const int G = T8 / 8;
        const long long total = (long long)N * OHh * OWw * G;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
            const int g = (int)(i % G);
            long long t = i / G;
            const int x = (int)(t % OWw), y = (int)((t / OWw) % OHh), n = (int)(t / ((long long)OWw * OHh));
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int tp = g * 8 + j;
                v[j] = 0.f;
                if (tp < taps) {
                    const int r = tp / k, s_ = tp - r * k;
                    if (mode == 0) {
                        v[j] = to_f(at<T>(big, ((size_t)n * BH + y + r) * BW + x + s_, 0)[0]);
                    } else {
                        const int py = y - r, px = x - s_;
                        if (py >= 0 && py < H && px >= 0 && px < W) v[j] = to_f(at<T>(small, ((size_t)n * H + py) * W + px, 0)[0]);
                    }
                }
            }
            if (mode == 0) Vec8<T>::store(at<T>(small, ((size_t)n * H + y) * W + x, g * 8), v);      // `small` view = the patches tensor
            else Vec8<T>::store(at<T>(big, ((size_t)n * BH + y) * BW + x, g * 8), v);
        }
    } else {
        // one thread = one pixel, sums over all taps into channel 0
        const int OHh = mode == 1 ? BH : H, OWw = mode == 1 ? BW : W;
        const long long total = (long long)N * OHh * OWw;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
            const int x = (int)(i % OWw), y = (int)((i / OWw) % OHh), n = (int)(i / ((long long)OWw * OHh));
            float acc = (mode == 2 && bias) ? bias[0] : 0.f;
            for (int tp = 0; tp < taps; ++tp) {
                const int r = tp / k, s_ = tp - r * k;
                if (mode == 2) {
                    acc += to_f(at<T>(big, ((size_t)n * BH + y + r) * BW + x + s_, tp)[0]);
                } else {
                    const int py = y - r, px = x - s_;
                    if (py >= 0 && py < H && px >= 0 && px < W) acc += to_f(at<T>(small, ((size_t)n * H + py) * W + px, tp)[0]);
                }
            }
            float v[8] = {acc, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (mode == 2) Vec8<T>::store(at<T>(small, ((size_t)n * H + y) * W + x, 0), v);
            else Vec8<T>::store(at<T>(big, ((size_t)n * BH + y) * BW + x, 0), v);
        }
    }
}

// virtual 1x1 kernels of the tap-folded 7x7 convs.  kind 0 (stem, Cin = 1): w1[t][co] <-> w[t][ci = 0][co];
// kind 1 (head, Cout = 1): w1[ci][t] <-> w[t][ci][co = 0].  dir 0 writes w1, dir 1 adds the gradient of w1 into w.
__global__ void tapfold_weights_kernel(float* __restrict__ w, int taps, int T8, int Cin, int Cout, float* __restrict__ w1, int kind, int dir) {
    const int rows = kind == 0 ? T8 : Cin, cols = kind == 0 ? Cout : T8;
    const long long total = (long long)rows * cols;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cols), r = (int)(i / cols);
        const int t = kind == 0 ? r : c;
        const bool valid = t < taps;
        const size_t m = kind == 0 ? ((size_t)t * Cin + 0) * Cout + c : ((size_t)t * Cin + r) * Cout + 0;
        if (dir == 0) w1[i] = valid ? w[m] : 0.f;
        else if (valid) w[m] += w1[i];
    }
}

// ---- overlapping tile grid of HelperFunctions.tile_image / stitch_image (:17-141) on the device --------------------------
// Tiles are numbered x-major (k = ix * ny + iy) like the reference's loops; xs / ys are the tile offsets along each axis.
// gather: tiles[k - k0][y][x] = img[ys[iy] + y][xs[ix] + x], zero outside the image.
__global__ void tile_gather_kernel(const float* __restrict__ img, int H, int W, float* __restrict__ tiles, int th, int tw,
                                   const int* __restrict__ xs, const int* __restrict__ ys, int ny, int k0, int count) {
    const long long total = (long long)count * th * tw;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % tw), y = (int)((i / tw) % th), k = k0 + (int)(i / ((long long)tw * th));
        const int gx = xs[k / ny] + x, gy = ys[k % ny] + y;
        tiles[i] = (gx < W && gy < H) ? img[(size_t)gy * W + gx] : 0.f;
    }
}

// stitch: one thread per output pixel walks the tiles in the reference's order.  mode 0: maximum (with the zero the
// reference's output buffer starts from), mode 1: average over the covering tiles, mode 2: every tile contributes its
// centre (half of the overlap cropped on each inner side); where two crops still meet, the later tile wins like the
// reference's sequential assignment.
__global__ void tile_stitch_kernel(const float* __restrict__ tiles, int th, int tw, float* __restrict__ out, int H, int W,
                                   const int* __restrict__ xs, int nx, const int* __restrict__ ys, int ny, int mode, int ovx, int ovy) {
    const long long total = (long long)H * W;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W), y = (int)(i / W);
        float acc = 0.f;
        int cnt = 0;
        for (int ix = 0; ix < nx; ++ix) {
            const int ox = xs[ix];
            int x_lo = ox, x_hi = min(ox + tw, W);
            if (mode == 2) { x_lo = ox + (ix == 0 ? 0 : ovx); x_hi = min(ox + tw - (ix == nx - 1 ? 0 : ovx), W); }
            if (x < x_lo || x >= x_hi) continue;
            for (int iy = 0; iy < ny; ++iy) {
                const int oy = ys[iy];
                int y_lo = oy, y_hi = min(oy + th, H);
                if (mode == 2) { y_lo = oy + (iy == 0 ? 0 : ovy); y_hi = min(oy + th - (iy == ny - 1 ? 0 : ovy), H); }
                if (y < y_lo || y >= y_hi) continue;
                const float v = tiles[((size_t)(ix * ny + iy) * th + (y - oy)) * tw + (x - ox)];
                if (mode == 0) acc = fmaxf(acc, v);
                else if (mode == 1) { acc += v; ++cnt; }
                else acc = v;
            }
        }
        // the reference divides by a uint8 count image: an uncovered pixel (cannot happen for a valid grid) would be 0/0
        out[i] = mode == 1 ? acc / (float)cnt : acc;
    }
}

// ---- parity mode on the tensor cores: fp32 tensors as split bf16 operands ------------------------------------------------
// x = xh + xm + xl exactly to 2^-24 |x| with xh = bf16(x), xm = bf16(x - xh), xl = bf16(x - xh - xm).
// 3 terms: x*w ~ xh*wh + xl'*wh + xh*wl'   (two-way split, relative error ~2^-16: measured 1e-3 on the UNet's sigmoid map
//          after 61 layers -- not enough for a bit-exact mask)
// 6 terms: x*w ~ xh*wh + xh*wm + xm*wh + xh*wl + xl*wh + xm*wm   (all products above 2^-24: fp32-grade)
// i.e. ONE tensor-core conv over the stacked operand [xh | xh | xm | xh | xl | xm] (6C channels) against the weights
// [wh ; wm ; wh ; wl ; wh ; wm] stacked along the contraction axis, accumulated in fp32.  (3 terms: [xh | xl' | xh] . [wh ; wh ; wl'].)
__device__ __forceinline__ void split3(float v, float& h, float& m, float& l) {
    h = __bfloat162float(__float2bfloat16_rn(v));
    const float r = v - h;
    m = __bfloat162float(__float2bfloat16_rn(r));
    l = r - m;
}

__global__ void split_bf16_kernel(PView src, PView dst, long long n_pixels, int C8, int C, int terms) {
    const long long total = n_pixels * C8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C8) * 8;
        const size_t px = (size_t)(i / C8);
        float v[8], hi[8], mid[8], lo[8];
        Vec8<float>::load(at<float>(src, px, c), v);
#pragma unroll
        for (int k = 0; k < 8; ++k) split3(v[k], hi[k], mid[k], lo[k]);
        if (terms == 3) {
#pragma unroll
            for (int k = 0; k < 8; ++k) mid[k] = v[k] - hi[k];           // two-way split: the whole residual
            Vec8<bf16>::store(at<bf16>(dst, px, c), hi);
            Vec8<bf16>::store(at<bf16>(dst, px, C + c), mid);
            Vec8<bf16>::store(at<bf16>(dst, px, 2 * C + c), hi);
        } else {
            Vec8<bf16>::store(at<bf16>(dst, px, c), hi);
            Vec8<bf16>::store(at<bf16>(dst, px, C + c), hi);
            Vec8<bf16>::store(at<bf16>(dst, px, 2 * C + c), mid);
            Vec8<bf16>::store(at<bf16>(dst, px, 3 * C + c), hi);
            Vec8<bf16>::store(at<bf16>(dst, px, 4 * C + c), lo);
            Vec8<bf16>::store(at<bf16>(dst, px, 5 * C + c), mid);
        }
    }
}

// w (R,S,Cin,Cout) fp32 -> the stacked fp32 kernel the packer then rounds to bf16 (block values are chosen so that their
// bf16 rounding is exactly wh / wm / wl).  axis 0: stacked along the input channels (forward); axis 1: along the output
// channels (the flipped pack then contracts over them: data gradient).
__global__ void split_weights_kernel(const float* __restrict__ w, int taps, int Cin, int Cout, float* __restrict__ ws, int axis, int terms) {
    const long long total = (long long)taps * Cin * Cout;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(i % Cout), ci = (int)((i / Cout) % Cin), t = (int)(i / ((long long)Cout * Cin));
        float h, m, l;
        split3(w[i], h, m, l);
        float blk[6];
        if (terms == 3) { blk[0] = h; blk[1] = h; blk[2] = w[i] - h; }
        else { blk[0] = h; blk[1] = m; blk[2] = h; blk[3] = l; blk[4] = h; blk[5] = m; }
        for (int b = 0; b < terms; ++b) {
            if (axis == 0) ws[((size_t)t * terms * Cin + (size_t)b * Cin + ci) * Cout + co] = blk[b];
            else ws[((size_t)t * Cin + ci) * terms * Cout + (size_t)b * Cout + co] = blk[b];
        }
    }
}

static inline int grid_for(long long total, int block = 256) {
    long long b = cdivl(total, block);
    const long long cap = 148LL * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace semb

using namespace semb;

extern "C" int semb_maxpool2x2_fwd(const semb_tensor* x, const semb_tensor* y, int32_t N, int32_t H, int32_t W,
                                   int32_t dtype, void* stream) {
    SEMB_REQUIRE(view_ok(x) && view_ok(y) && x->C == y->C, SEMB_EALIGN, "maxpool fwd: bad views");
    SEMB_REQUIRE(N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, SEMB_ESHAPE, "maxpool fwd: H,W must be even");
    const int C8 = x->C / 8;
    const long long total = (long long)N * (H / 2) * (W / 2) * C8;
    if (dtype == SEMB_BF16) launch_pdl(maxpool_kernel<bf16, false>, dim3(grid_for(total)), dim3(256), 0, as_stream(stream), pv(x), pv(y), PView{}, N, H, W, C8, 0);
    else launch_pdl(maxpool_kernel<float, false>, dim3(grid_for(total)), dim3(256), 0, as_stream(stream), pv(x), pv(y), PView{}, N, H, W, C8, 0);
    return check_launch("maxpool_fwd");
}

extern "C" int semb_maxpool2x2_bwd(const semb_tensor* x, const semb_tensor* dy, const semb_tensor* dx, int32_t N, int32_t H,
                                   int32_t W, int32_t dtype, int32_t accumulate, void* stream) {
    SEMB_REQUIRE(view_ok(x) && view_ok(dy) && view_ok(dx) && x->C == dy->C && x->C == dx->C, SEMB_EALIGN, "maxpool bwd: bad views");
    SEMB_REQUIRE(N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, SEMB_ESHAPE, "maxpool bwd: H,W must be even");
    const int C8 = x->C / 8;
    const long long total = (long long)N * (H / 2) * (W / 2) * C8;
    if (dtype == SEMB_BF16) launch_pdl(maxpool_kernel<bf16, true>, dim3(grid_for(total)), dim3(256), 0, as_stream(stream), pv(x), pv(dy), pv(dx), N, H, W, C8, accumulate);
    else launch_pdl(maxpool_kernel<float, true>, dim3(grid_for(total)), dim3(256), 0, as_stream(stream), pv(x), pv(dy), pv(dx), N, H, W, C8, accumulate);
    return check_launch("maxpool_bwd");
}

extern "C" int semb_pad_crop(const semb_tensor* x, const semb_tensor* y, int32_t N, int32_t H, int32_t W, int32_t OH,
                             int32_t OW, int32_t top, int32_t left, int32_t mode, int32_t dtype, int32_t accumulate,
                             void* stream) {
    SEMB_REQUIRE(view_ok(x) && view_ok(y) && x->C == y->C, SEMB_EALIGN, "pad_crop: bad views");
    SEMB_REQUIRE(N > 0 && H > 0 && W > 0 && OH > 0 && OW > 0 && top >= 0 && left >= 0 && mode >= 0 && mode <= 3, SEMB_ESHAPE,
                 "pad_crop: bad geometry");
    if (mode == 0 || mode == 2) {
        SEMB_REQUIRE(OH >= H + top && OW >= W + left, SEMB_ESHAPE, "pad_crop: output smaller than padded input");
        if (mode == 0) SEMB_REQUIRE(top < H && left < W && OH - H - top < H && OW - W - left < W, SEMB_ESHAPE, "pad_crop: reflect pad wider than image");
    } else {
        SEMB_REQUIRE(H >= OH + top && W >= OW + left, SEMB_ESHAPE, "pad_crop: crop window outside the input");
    }
    const int C8 = x->C / 8;
    const long long total = (long long)N * OH * OW * C8;
    if (dtype == SEMB_BF16) launch_pdl(pad_crop_kernel<bf16>, dim3(grid_for(total)), dim3(256), 0, as_stream(stream), pv(x), pv(y), N, H, W, OH, OW, top, left, mode, C8, accumulate);
    else launch_pdl(pad_crop_kernel<float>, dim3(grid_for(total)), dim3(256), 0, as_stream(stream), pv(x), pv(y), N, H, W, OH, OW, top, left, mode, C8, accumulate);
    return check_launch("pad_crop");
}

extern "C" int semb_loss_wbce(const semb_tensor* p, const float* y_true, const semb_tensor* dp, int64_t count,
                              float weighting, float* out, int32_t dtype, void* stream) {
    SEMB_REQUIRE(view_ok(p) && y_true && out && count > 0 && (!dp || view_ok(dp)), SEMB_ESHAPE, "loss_wbce: bad arguments");
    if (dtype == SEMB_BF16) loss_wbce_kernel<bf16><<<grid_for(count), 256, 0, as_stream(stream)>>>(pv(p), y_true, pv(dp), count, weighting, out);
    else loss_wbce_kernel<float><<<grid_for(count), 256, 0, as_stream(stream)>>>(pv(p), y_true, pv(dp), count, weighting, out);
    return check_launch("loss_wbce");
}

extern "C" int semb_loss_wbce_logits(const semb_tensor* z, const float* scale, const float* shift, const float* y_true,
                                     const semb_tensor* dp, int64_t count, float weighting, float* out, int32_t dtype, void* stream) {
    SEMB_REQUIRE(view_ok(z) && scale && shift && y_true && out && count > 0 && (!dp || view_ok(dp)), SEMB_ESHAPE,
                 "loss_wbce_logits: bad arguments");
    if (dtype == SEMB_BF16) launch_pdl(loss_wbce_logits_kernel<bf16>, dim3(grid_for(count)), dim3(256), 0, as_stream(stream), pv(z), scale, shift, y_true, pv(dp), count, weighting, out);
    else launch_pdl(loss_wbce_logits_kernel<float>, dim3(grid_for(count)), dim3(256), 0, as_stream(stream), pv(z), scale, shift, y_true, pv(dp), count, weighting, out);
    return check_launch("loss_wbce_logits");
}

extern "C" int semb_loss_l1_l2(const semb_tensor* a, const semb_tensor* b, float target, int32_t kind, int64_t n_pixels,
                               int32_t c_logical, float gscale, const semb_tensor* da, int32_t accumulate, float* out,
                               int32_t dtype, void* stream) {
    SEMB_REQUIRE(view_ok(a) && (!b || (view_ok(b) && b->C == a->C)) && (!da || (view_ok(da) && da->C == a->C)) && out && n_pixels > 0,
                 SEMB_ESHAPE, "loss_l1_l2: bad arguments");
    SEMB_REQUIRE(kind == 0 || kind == 1, SEMB_ESHAPE, "loss_l1_l2: kind must be 0 (L1) or 1 (L2)");
    const int C8 = a->C / 8;
    const long long total = n_pixels * C8;
    if (dtype == SEMB_BF16) launch_pdl(loss_l1_l2_kernel<bf16>, dim3(grid_for(total)), dim3(256), 0, as_stream(stream), pv(a), pv(b), target, kind, n_pixels, C8, c_logical, gscale, pv(da), accumulate, out);
    else launch_pdl(loss_l1_l2_kernel<float>, dim3(grid_for(total)), dim3(256), 0, as_stream(stream), pv(a), pv(b), target, kind, n_pixels, C8, c_logical, gscale, pv(da), accumulate, out);
    return check_launch("loss_l1_l2");
}

extern "C" int semb_adam_step(float* w, const float* g, float* m, float* v, int64_t n, const float* lr_ptr, float beta1,
                              float beta2, float eps, float gscale, void* state, void* stream) {
    SEMB_REQUIRE(w && g && m && v && lr_ptr && state && n > 0, SEMB_ESHAPE, "adam: bad arguments");
    SEMB_REQUIRE(((uintptr_t)w % 16) == 0 && ((uintptr_t)g % 16) == 0 && ((uintptr_t)m % 16) == 0 && ((uintptr_t)v % 16) == 0,
                 SEMB_EALIGN, "adam: buffers must be 16B aligned");
    launch_pdl(adam_tick_kernel, dim3(1), dim3(1), 0, as_stream(stream), reinterpret_cast<AdamState*>(state), lr_ptr, beta1, beta2);
    int rc = check_launch("adam_tick");
    if (rc) return rc;
    launch_pdl(adam_kernel, dim3(grid_for(cdivl(n, 4))), dim3(256), 0, as_stream(stream), w, g, m, v, n, reinterpret_cast<const AdamState*>(state), beta1, beta2, eps, gscale);
    return check_launch("adam");
}

extern "C" int semb_fill_f32(float* p, int64_t n, float value, void* stream) {
    SEMB_REQUIRE(p && n >= 0, SEMB_ESHAPE, "fill: bad arguments");
    if (n == 0) return SEMB_OK;
    launch_pdl(fill_kernel, dim3(grid_for(n)), dim3(256), 0, as_stream(stream), p, n, value);
    return check_launch("fill");
}

extern "C" int semb_cast_in(const float* src, int32_t src_C, const semb_tensor* dst, int64_t n_pixels, int32_t dtype, void* stream) {
    SEMB_REQUIRE(src && view_ok(dst) && n_pixels > 0 && src_C > 0 && src_C <= dst->C, SEMB_ESHAPE, "cast_in: bad arguments");
    const int C8 = dst->C / 8;
    if (dtype == SEMB_BF16) launch_pdl(cast_in_kernel<bf16>, dim3(grid_for(n_pixels * C8)), dim3(256), 0, as_stream(stream), src, src_C, pv(dst), n_pixels, C8);
    else launch_pdl(cast_in_kernel<float>, dim3(grid_for(n_pixels * C8)), dim3(256), 0, as_stream(stream), src, src_C, pv(dst), n_pixels, C8);
    return check_launch("cast_in");
}

extern "C" int semb_cast_out(const semb_tensor* src, float* dst, int32_t dst_C, int64_t n_pixels, int32_t dtype, void* stream) {
    SEMB_REQUIRE(dst && view_ok(src) && n_pixels > 0 && dst_C > 0 && dst_C <= src->C, SEMB_ESHAPE, "cast_out: bad arguments");
    const int C8 = src->C / 8;
    if (dtype == SEMB_BF16) cast_out_kernel<bf16><<<grid_for(n_pixels * C8), 256, 0, as_stream(stream)>>>(pv(src), dst, dst_C, n_pixels, C8);
    else cast_out_kernel<float><<<grid_for(n_pixels * C8), 256, 0, as_stream(stream)>>>(pv(src), dst, dst_C, n_pixels, C8);
    return check_launch("cast_out");
}

extern "C" int semb_pixel_shuffle2x(const semb_tensor* src, const semb_tensor* dst, int32_t N, int32_t H, int32_t W, int32_t DH,
                                    int32_t DW, const float* bias, int32_t dir, int32_t acc, int32_t dtype, void* stream) {
    SEMB_REQUIRE(view_ok(src) && view_ok(dst) && src->C == 4 * dst->C, SEMB_ESHAPE, "pixel_shuffle2: src must have 4x the channels of dst");
    SEMB_REQUIRE(N > 0 && H > 0 && W > 0 && (dir == 0 || dir == 1) && DH <= 2 * H && DW <= 2 * W && DH >= 2 * H - 1 && DW >= 2 * W - 1,
                 SEMB_ESHAPE, "pixel_shuffle2: bad geometry");
    const int C8 = dst->C / 8;
    const long long total = (long long)N * H * W * 4 * C8;
    if (dtype == SEMB_BF16) launch_pdl(pixel_shuffle2_kernel<bf16>, dim3(grid_for(total)), dim3(256), 0, as_stream(stream), pv(src), pv(dst), N, H, W, DH, DW, C8, bias, dir, acc);
    else launch_pdl(pixel_shuffle2_kernel<float>, dim3(grid_for(total)), dim3(256), 0, as_stream(stream), pv(src), pv(dst), N, H, W, DH, DW, C8, bias, dir, acc);
    return check_launch("pixel_shuffle2");
}

extern "C" int semb_pixel_shuffle2(const semb_tensor* src, const semb_tensor* dst, int32_t N, int32_t H, int32_t W,
                                   const float* bias, int32_t dir, int32_t dtype, void* stream) {
    return semb_pixel_shuffle2x(src, dst, N, H, W, 2 * H, 2 * W, bias, dir, 0, dtype, stream);
}

extern "C" int semb_split_bf16(const semb_tensor* src, const semb_tensor* dst, int64_t n_pixels, void* stream) {
    SEMB_REQUIRE(src && dst && src->ptr && dst->ptr && n_pixels > 0 && src->C > 0 && (src->C % 8) == 0 &&
                 (dst->C == 3 * src->C || dst->C == 6 * src->C) && (src->pitch % 8) == 0 && (src->coff % 8) == 0 && view_ok(dst), SEMB_ESHAPE,
                 "split_bf16: dst must have 3x or 6x the channels of src");
    const int C8 = src->C / 8;
    split_bf16_kernel<<<grid_for(n_pixels * C8), 256, 0, as_stream(stream)>>>(pv(src), pv(dst), n_pixels, C8, src->C, dst->C / src->C);
    return check_launch("split_bf16");
}

extern "C" int semb_split_weights(const float* w, int32_t R, int32_t S, int32_t Cin, int32_t Cout, float* ws, int32_t axis, int32_t terms,
                                  void* stream) {
    SEMB_REQUIRE(w && ws && R > 0 && S > 0 && Cin > 0 && Cout > 0 && (axis == 0 || axis == 1) && (terms == 3 || terms == 6), SEMB_ESHAPE,
                 "split_weights: bad arguments");
    split_weights_kernel<<<grid_for((long long)R * S * Cin * Cout), 256, 0, as_stream(stream)>>>(w, R * S, Cin, Cout, ws, axis, terms);
    return check_launch("split_weights");
}

extern "C" int semb_tile_gather(const float* img, int32_t H, int32_t W, float* tiles, int32_t th, int32_t tw, const int32_t* xs, int32_t nx,
                                const int32_t* ys, int32_t ny, int32_t k0, int32_t count, void* stream) {
    SEMB_REQUIRE(img && tiles && xs && ys && H > 0 && W > 0 && th > 0 && tw > 0 && nx > 0 && ny > 0 && k0 >= 0 && count > 0 &&
                 k0 + count <= nx * ny, SEMB_ESHAPE, "tile_gather: bad arguments");
    tile_gather_kernel<<<grid_for((long long)count * th * tw), 256, 0, as_stream(stream)>>>(img, H, W, tiles, th, tw, xs, ys, ny, k0, count);
    return check_launch("tile_gather");
}

extern "C" int semb_tile_stitch(const float* tiles, int32_t th, int32_t tw, float* out, int32_t H, int32_t W, const int32_t* xs, int32_t nx,
                                const int32_t* ys, int32_t ny, int32_t mode, void* stream) {
    SEMB_REQUIRE(tiles && out && xs && ys && H > 0 && W > 0 && th > 0 && tw > 0 && nx > 0 && ny > 0 && mode >= 0 && mode <= 2, SEMB_ESHAPE,
                 "tile_stitch: bad arguments");
    // half of the overlap that mode 2 crops from each inner side (HelperFunctions.py:84-85)
    const int ovx = nx > 1 ? (tw * nx - W) / (2 * (nx - 1)) : 0;
    const int ovy = ny > 1 ? (th * ny - H) / (2 * (ny - 1)) : 0;
    tile_stitch_kernel<<<grid_for((long long)H * W), 256, 0, as_stream(stream)>>>(tiles, th, tw, out, H, W, xs, nx, ys, ny, mode, ovx, ovy);
    return check_launch("tile_stitch");
}

extern "C" int semb_upsample2x(const semb_tensor* small, const semb_tensor* big, int32_t N, int32_t H, int32_t W, int32_t dir,
                               int32_t acc, int32_t dtype, void* stream) {
    SEMB_REQUIRE(view_ok(small) && view_ok(big) && small->C == big->C && N > 0 && H > 0 && W > 0 && (dir == 0 || dir == 1), SEMB_ESHAPE,
                 "upsample2x: bad arguments");
    const int C8 = small->C / 8;
    const long long total = (long long)N * H * W * C8;
    if (dtype == SEMB_BF16) upsample2x_kernel<bf16><<<grid_for(total), 256, 0, as_stream(stream)>>>(pv(small), pv(big), N, H, W, C8, dir, acc);
    else upsample2x_kernel<float><<<grid_for(total), 256, 0, as_stream(stream)>>>(pv(small), pv(big), N, H, W, C8, dir, acc);
    return check_launch("upsample2x");
}

extern "C" int semb_mask_mul(const semb_tensor* x, const semb_tensor* z, const semb_tensor* m, const semb_tensor* y, int64_t n_pixels,
                             float negative_slope, int32_t accumulate, int32_t dtype, void* stream) {
    SEMB_REQUIRE(view_ok(x) && view_ok(y) && x->C == y->C && n_pixels > 0 && (!z || (view_ok(z) && z->C == x->C)) &&
                 (!m || (view_ok(m) && m->C == x->C)), SEMB_ESHAPE, "mask_mul: bad arguments");
    const int C8 = x->C / 8;
    const PView zv = z ? pv(z) : pv(x), mv = m ? pv(m) : pv(x);
    if (dtype == SEMB_BF16) launch_pdl(mask_mul_kernel<bf16>, dim3(grid_for(n_pixels * C8)), dim3(256), 0, as_stream(stream), pv(x), zv, mv, pv(y),
                                       (long long)n_pixels, C8, negative_slope, z ? 1 : 0, m ? 1 : 0, accumulate);
    else launch_pdl(mask_mul_kernel<float>, dim3(grid_for(n_pixels * C8)), dim3(256), 0, as_stream(stream), pv(x), zv, mv, pv(y),
                    (long long)n_pixels, C8, negative_slope, z ? 1 : 0, m ? 1 : 0, accumulate);
    return check_launch("mask_mul");
}

extern "C" int semb_gp_direction(const semb_tensor* g, const semb_tensor* u, int32_t N, int64_t HW, float scale, float* sums, int32_t dtype,
                                 void* stream) {
    SEMB_REQUIRE(view_ok(g) && view_ok(u) && g->C == u->C && N > 0 && HW > 0 && sums, SEMB_ESHAPE, "gp_direction: bad arguments");
    const int C8 = g->C / 8;
    if (dtype == SEMB_BF16) gp_direction_kernel<bf16><<<N, 256, 0, as_stream(stream)>>>(pv(g), pv(u), (long long)HW, C8, scale, sums);
    else gp_direction_kernel<float><<<N, 256, 0, as_stream(stream)>>>(pv(g), pv(u), (long long)HW, C8, scale, sums);
    return check_launch("gp_direction");
}

extern "C" int semb_merge_weights(float* wa, float* ws, int32_t Cin, int32_t Ca, int32_t Cs, float* w3, int32_t dir, void* stream) {
    SEMB_REQUIRE(wa && ws && w3 && Cin > 0 && Ca > 0 && Cs > 0 && (dir == 0 || dir == 1), SEMB_ESHAPE, "merge_weights: bad arguments");
    const long long total = 9LL * Cin * (Ca + Cs);
    merge_weights_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(wa, ws, Cin, Ca, Cs, w3, dir);
    return check_launch("merge_weights");
}

extern "C" int semb_s2d_weights(float* w, int32_t k, int32_t pad_t, int32_t pad_l, int32_t Cin, int32_t Cout, float* w3, int32_t dir,
                                void* stream) {
    SEMB_REQUIRE(w && w3 && (k == 3 || k == 4 || (k == 5 && pad_t == 1 && pad_l == 1)) && (pad_t == 0 || pad_t == 1) && (pad_l == 0 || pad_l == 1) && Cin > 0 && Cout > 0 &&
                 (dir == 0 || dir == 1), SEMB_ESHAPE, "s2d_weights: bad arguments");
    const long long total = 9LL * 4 * Cin * Cout;
    s2d_weights_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(w, k, pad_t, pad_l, Cin, Cout, w3, dir);
    return check_launch("s2d_weights");
}

extern "C" int semb_fold_stats4(const void* temp, void* stats, int32_t groups, int32_t C, int32_t stats_nstride, int32_t stats_cstride,
                                void* stream) {
    SEMB_REQUIRE(temp && stats && groups > 0 && C > 0, SEMB_ESHAPE, "fold_stats4: bad arguments");
    const int n = groups * 2 * C;
    fold_stats4_kernel<<<cdiv(n, 128), 128, 0, as_stream(stream)>>>(reinterpret_cast<const double*>(temp), reinterpret_cast<double*>(stats), groups,
                                                                    C, stats_nstride, stats_cstride);
    return check_launch("fold_stats4");
}

extern "C" int semb_tap_patch(const semb_tensor* small, const semb_tensor* big, int32_t N, int32_t H, int32_t W, int32_t k,
                              const float* bias, int32_t mode, int32_t dtype, void* stream) {
    SEMB_REQUIRE(view_ok(small) && view_ok(big) && N > 0 && H > 0 && W > 0 && k > 1 && mode >= 0 && mode <= 3, SEMB_ESHAPE, "tap_patch: bad arguments");
    const int T8 = (k * k + 7) / 8 * 8;
    // the tap-channel tensor is `small` in modes 0/1 (patches) and `big` in modes 2/3 (per-tap conv on the padded domain)
    SEMB_REQUIRE((mode <= 1 ? small->C : big->C) == T8, SEMB_ESHAPE, "tap_patch: the tap tensor needs %d channels", T8);
    const long long total = mode == 0 ? (long long)N * H * W * (T8 / 8)
                          : mode == 3 ? (long long)N * (H + k - 1) * (W + k - 1) * (T8 / 8)
                          : mode == 1 ? (long long)N * (H + k - 1) * (W + k - 1) : (long long)N * H * W;
    if (dtype == SEMB_BF16) tap_patch_kernel<bf16><<<grid_for(total), 256, 0, as_stream(stream)>>>(pv(small), pv(big), N, H, W, k, T8, bias, mode);
    else tap_patch_kernel<float><<<grid_for(total), 256, 0, as_stream(stream)>>>(pv(small), pv(big), N, H, W, k, T8, bias, mode);
    return check_launch("tap_patch");
}

extern "C" int semb_tapfold_weights(float* w, int32_t k, int32_t Cin, int32_t Cout, float* w1, int32_t kind, int32_t dir, void* stream) {
    SEMB_REQUIRE(w && w1 && k > 1 && Cin > 0 && Cout > 0 && (kind == 0 || kind == 1) && (dir == 0 || dir == 1), SEMB_ESHAPE,
                 "tapfold_weights: bad arguments");
    const int T8 = (k * k + 7) / 8 * 8;
    const long long total = (long long)T8 * (kind == 0 ? Cout : Cin);
    tapfold_weights_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(w, k * k, T8, Cin, Cout, w1, kind, dir);
    return check_launch("tapfold_weights");
}
