// Fused normalisation / activation / residual kernels (HBM-bound, 128-bit accesses).
//
// Every tensor is NHWC with 8-padded channels, so one thread always moves 8 channels of one pixel
// (16 B of bf16, 32 B of fp32).  A block owns a contiguous range of pixels of ONE sample and all
// channels; per-channel reductions (BatchNorm / InstanceNorm moments and their backward sums) are
// carried in registers across the block's pixel loop, combined in shared memory and flushed with one
// atomic per channel per block.
#include "common.cuh"
#include <stdlib.h>

// pixels a thread keeps in flight per loop iteration (16-byte loads per operand), tuned on B200
#ifndef SEMB_AFF_U_FWD
#define SEMB_AFF_U_FWD 4
#endif
#ifndef SEMB_AFF_U_FWD_B
#define SEMB_AFF_U_FWD_B 2
#endif
#ifndef SEMB_AFF_U_BWD
#define SEMB_AFF_U_BWD 2
#endif

namespace semb {

struct View { const void* ptr; int pitch, coff; };
static inline View mkview(const semb_tensor* t) { return t ? View{t->ptr, t->pitch, t->coff} : View{nullptr, 0, 0}; }

template <typename T>
__device__ __forceinline__ const T* vptr(const View& v, long long pixel, int c) {
    return reinterpret_cast<const T*>(v.ptr) + (size_t)pixel * v.pitch + v.coff + c;
}
template <typename T>
__device__ __forceinline__ T* vptr_mut(const View& v, long long pixel, int c) {
    return reinterpret_cast<T*>(const_cast<void*>(v.ptr)) + (size_t)pixel * v.pitch + v.coff + c;
}

// Block geometry shared by all kernels of this file: C8 = C/8 channel lanes, rows = 256 / C8 pixel rows.
struct Lanes {
    int c8, rows, cg, prow;
    bool active;
    __device__ Lanes(int C) {
        c8 = C >> 3;
        rows = 256 / c8;
        cg = threadIdx.x % c8;
        prow = threadIdx.x / c8;
        active = prow < rows;
    }
};

struct AffArgs {
    int N, HW, C, act, actb, mode_a, mode_b, aff_nstride;
    View a, b, y, dy, da, db;
    const float *scale_a, *shift_a, *scale_b, *shift_b;
    const float *mean_a, *invstd_a, *c1_a, *c2_a, *mean_b, *invstd_b, *c1_b, *c2_b;
    float* stats; double* dstats; int stats_nstride, stats_cstride;
    int acc_a, acc_b;
    int ppb;  // pixels per block
    // finalize of the normalisations folded into the forward kernel (stats == NULL: scale / shift are read)
    semb_norm_fin fa, fb;
    // fused backward (reduce -> grid barrier -> apply) and apply-from-sums
    float count_a, count_b;
    float *dgamma_a, *dbeta_a, *dgamma_b, *dbeta_b;
    unsigned int* barrier;
    // b as a concatenation of up to three compact tensors (semb_affine_desc.nseg_b)
    int nseg_b; int seg_c0[3]; View seg_b[3], seg_db[3];
    // bf16 operands through a per-thread cp.async ring in dynamic shared memory (byte offset ring_off), see OperandRing
    int ring, ring_off;
    int rev;    // gradient pass: walk the block's pixel range backwards (what the reduction pass read last is still in L2)
};

// ---- per-thread asynchronous operand ring (bf16 tensors) ----------------------------------------------------------------
// ncu, round 1: the affine kernels reached 64-66 % of the DRAM peak with 35 % of the warps resident and 21-31 % issue
// utilisation -- latency-bound: with register prefetch a thread holds 2 pixels x 2-3 operands x 16 B in flight (49 KB per
// SM, ~7 MB on the chip, about the bandwidth-delay product with nothing to spare while a warp computes and stores).
// Here every thread copies ITS OWN 16-byte operand chunks S iterations ahead with cp.async (LDGSTS, L1 bypass) into a slot
// only it reads back: no block-level synchronisation, no registers held by loads in flight, 75-170 KB in flight per SM.
// Slot of (stage, operand) for thread t: ring + ((stage * NOPS + operand) * 256 + t) * 16 -- conflict-free LDS.128.
// p.ring = number of stages S (a power of two <= 8; 0 = register prefetch).  Measured on B200 (profiles/r02_ab_ring.txt): the
// ring lifts the two-operand forward from 3.8-4.3 to 5.2-5.7 TB/s and the reduction pass from 4.4-5.2 to 5.3-6.4 TB/s
// (97 % of the measured copy peak at C = 32); the gradient pass and the one-operand forward, already at 5.3-6.0 TB/s alone,
// gain nothing in isolation but the whole step is fastest with every kernel on the ring (12.45 vs 13.0 ms, r02_ab_ring3.txt).
__device__ __forceinline__ void cp_async_wait_stages(int S) {      // wait until at most S - 1 groups are pending
    switch (S) {
        case 8: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
        case 4: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    }
}
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}

// The view and the channel offset inside it that hold channel c of the b operand (resp. of its gradient).
__device__ __forceinline__ View b_view(const AffArgs& p, int c, int& cc, bool grad) {
    cc = c;
    if (p.nseg_b == 0) return grad ? p.db : p.b;
    int s = 0;
    if (p.nseg_b > 1 && c >= p.seg_c0[1]) s = 1;
    if (p.nseg_b > 2 && c >= p.seg_c0[2]) s = 2;
    cc = c - p.seg_c0[s];
    return grad ? p.seg_db[s] : p.seg_b[s];
}

__device__ __forceinline__ float act_bwd_from_u(float u, int act) {
    switch (act) {
        case SEMB_ACT_RELU: return u > 0.f ? 1.f : 0.f;
        case SEMB_ACT_LEAKY: return u > 0.f ? 1.f : 0.2f;
        case SEMB_ACT_SIGMOID: { const float s = 1.f / (1.f + expf(-u)); return s * (1.f - s); }
        case SEMB_ACT_TANH: { const float t = tanhf(u); return 1.f - t * t; }
        default: return 1.f;
    }
}

// Raw (unconverted) 8-channel loads: all loads of an iteration are issued before any value is unpacked, and a bf16
// vector waits in 4 registers instead of 8.
template <typename T> struct Raw8;
template <> struct Raw8<bf16> {
    uint4 r;
    __device__ __forceinline__ void load(const bf16* p) { r = *reinterpret_cast<const uint4*>(p); }
    __device__ __forceinline__ void unpack(float (&v)[8]) const {
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
    }
};
template <> struct Raw8<float> {
    float4 a, b;
    __device__ __forceinline__ void load(const float* p) { a = *reinterpret_cast<const float4*>(p); b = *reinterpret_cast<const float4*>(p + 4); }
    __device__ __forceinline__ void unpack(float (&v)[8]) const {
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
};

// Block-level column sums of per-thread-row partials sm[rows][ncol] with all 256 threads: G = 256/ncol row groups are
// summed in parallel into sm2[G][ncol], then `emit(col, total)` is called once per column.  Fixed order: deterministic.
template <typename F>
__device__ __forceinline__ void block_colsum(const float* sm, float* sm2, int rows, int ncol, F emit) {
    int G = 256 / ncol;                 // sm2 holds G * ncol <= 256 floats
    if (G > rows) G = rows;
    __syncthreads();
    if (G <= 1) {
        for (int col = threadIdx.x; col < ncol; col += 256) {
            float t = 0.f;
            for (int r = 0; r < rows; ++r) t += sm[r * ncol + col];
            emit(col, t);
        }
        return;
    }
    for (int idx = threadIdx.x; idx < ncol * G; idx += 256) {
        const int col = idx % ncol, gi = idx / ncol;
        float t = 0.f;
        for (int r = gi; r < rows; r += G) t += sm[r * ncol + col];
        sm2[idx] = t;
    }
    __syncthreads();
    for (int col = threadIdx.x; col < ncol; col += 256) {
        float t = 0.f;
        for (int gi = 0; gi < G; ++gi) t += sm2[gi * ncol + col];
        emit(col, t);
    }
}

template <int K>
__device__ __forceinline__ void ld_params(const float* p, size_t off, float (&v)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = p[off + i];
}

// Folded semb_norm_finalize: the thread's 8 channels of group g, same arithmetic as norm_finalize_kernel.  `publish`:
// this thread also writes scale / shift / mean / invstd (read by the backward kernels) and, for g == 0, the moving
// statistics.
__device__ __forceinline__ void fin_params(const semb_norm_fin& f, int g, int c, bool publish, bool moving, float (&sc)[8], float (&sh)[8]) {
    const double* st = reinterpret_cast<const double*>(f.stats) + (size_t)g * f.stats_nstride;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const double dm = st[c + i] / (double)f.count;
        const float mean = (float)dm;
        const float var = (float)(st[f.cstride + c + i] / (double)f.count - dm * dm);
        const float inv = rsqrtf(var + f.eps);
        const float ga = f.gamma ? f.gamma[c + i] : 1.f;
        sc[i] = ga * inv;
        sh[i] = f.beta[c + i] - mean * sc[i];
        if (publish) {
            const size_t o = (size_t)g * f.cstride + c + i;
            f.scale[o] = sc[i];
            f.shift[o] = sh[i];
            if (f.mean) f.mean[o] = mean;
            if (f.invstd) f.invstd[o] = inv;
            if (moving && f.moving_mean) {
                f.moving_mean[c + i] = f.moving_mean[c + i] * f.momentum + mean * (1.f - f.momentum);
                f.moving_var[c + i] = f.moving_var[c + i] * f.momentum + var * (1.f - f.momentum);
            }
        }
    }
}

// y = act(a*sa+ta [+ actb(b*sb+tb)]), optional fp64 moments of y.  U pixels per thread are loaded before any is used.
template <typename T, bool HAS_B, int ACT, int ACTB>
__global__ void __launch_bounds__(256, HAS_B ? 2 : 3) affine_act_fwd_kernel(const AffArgs p) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float sm[];  // [rows][2][C] + [<=256] when moments are requested
    constexpr int U = HAS_B ? SEMB_AFF_U_FWD_B : SEMB_AFF_U_FWD;
    const int act = ACT >= 0 ? ACT : p.act, actb = ACTB >= 0 ? ACTB : p.actb;
    const Lanes L(p.C);
    const int n = blockIdx.y;
    const int c = L.cg * 8;
    int cb = c;
    const View bv = b_view(p, c, cb, false);
    const long long pix0 = (long long)n * p.HW;
    const int begin = blockIdx.x * p.ppb, end = min(p.HW, begin + p.ppb);
    const size_t aoff = (size_t)n * p.aff_nstride + c;

    float sa[8], ta[8], sb[8], tb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { sa[i] = 1.f; ta[i] = 0.f; sb[i] = 1.f; tb[i] = 0.f; }
    if (L.active) {
        const int g = p.aff_nstride != 0 ? n : 0;
        const bool publish = blockIdx.x == 0 && L.prow == 0 && (p.aff_nstride != 0 || n == 0);
        const bool moving = publish && n == 0 && p.aff_nstride == 0;
        if (p.mode_a != SEMB_AFF_NONE) {
            if (p.fa.stats) fin_params(p.fa, g, c, publish, moving, sa, ta);
            else { ld_params<0>(p.scale_a, aoff, sa); ld_params<0>(p.shift_a, aoff, ta); }
        }
        if (HAS_B && p.mode_b != SEMB_AFF_NONE) {
            if (p.fb.stats) fin_params(p.fb, g, c, publish, moving, sb, tb);
            else { ld_params<0>(p.scale_b, aoff, sb); ld_params<0>(p.shift_b, aoff, tb); }
        }
    }
    float s1[8], s2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s1[i] = 0.f; s2[i] = 0.f; }

    bool ringed = false;
    if constexpr (sizeof(T) == 2) {
        if (p.ring && L.active) {
            ringed = true;
            const int S = p.ring; constexpr int NOPS = HAS_B ? 2 : 1;
            const uint32_t ring = smem_addr(sm) + (uint32_t)p.ring_off + threadIdx.x * 16u;
            const int first = begin + L.prow;
            const int niter = first < end ? (end - first + L.rows - 1) / L.rows : 0;
            auto issue = [&](int it) {
                if (it < niter) {
                    const long long px = pix0 + first + (long long)it * L.rows;
                    const uint32_t slot = ring + (uint32_t)((it & (S - 1)) * NOPS) * 4096u;
                    cp_async16(slot, vptr<T>(p.a, px, c));
                    if (HAS_B) cp_async16(slot + 4096u, vptr<T>(bv, px, cb));
                }
                cp_async_commit();          // empty groups keep the group count uniform
            };
            for (int i = 0; i < S - 1; ++i) issue(i);
            for (int it = 0; it < niter; ++it) {
                issue(it + S - 1);
                cp_async_wait_stages(S);
                const uint32_t slot = ring + (uint32_t)((it & (S - 1)) * NOPS) * 4096u;
                Raw8<T> ra, rb;
                ra.r = lds128(slot);
                if (HAS_B) rb.r = lds128(slot + 4096u);
                float va[8], vb[8], vy[8];
                ra.unpack(va);
                if (HAS_B) rb.unpack(vb);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float t = fmaf(va[i], sa[i], ta[i]);
                    if (HAS_B) t += act_fwd(fmaf(vb[i], sb[i], tb[i]), actb);
                    vy[i] = act_fwd(t, act);
                    s1[i] += vy[i];
                    s2[i] += vy[i] * vy[i];
                }
                Vec8<T>::store(vptr_mut<T>(p.y, pix0 + first + (long long)it * L.rows, c), vy);
            }
        }
    }
    if (L.active && !ringed) {
        for (int base = begin + L.prow; base < end; base += L.rows * U) {
            Raw8<T> ra[U], rb[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int px = base + u * L.rows;
                if (px < end) {
                    ra[u].load(vptr<T>(p.a, pix0 + px, c));
                    if (HAS_B) rb[u].load(vptr<T>(bv, pix0 + px, cb));
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int px = base + u * L.rows;
                if (px < end) {
                    float va[8], vb[8], vy[8];
                    ra[u].unpack(va);
                    if (HAS_B) rb[u].unpack(vb);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float t = fmaf(va[i], sa[i], ta[i]);
                        if (HAS_B) t += act_fwd(fmaf(vb[i], sb[i], tb[i]), actb);
                        vy[i] = act_fwd(t, act);
                        s1[i] += vy[i];
                        s2[i] += vy[i] * vy[i];
                    }
                    Vec8<T>::store(vptr_mut<T>(p.y, pix0 + px, c), vy);
                }
            }
        }
    }
    if (p.dstats) {
        // deterministic: partials [rows][2][C] in shared memory, fixed-order column sums, fp64 atomics across blocks
        if (L.active) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                sm[(L.prow * 2) * p.C + c + i] = s1[i];
                sm[(L.prow * 2 + 1) * p.C + c + i] = s2[i];
            }
        }
        float* sm2 = sm + L.rows * 2 * p.C;
        block_colsum(sm, sm2, L.rows, 2 * p.C, [&](int col, float t) {
            const int k = col / p.C, i = col - k * p.C;
            atomicAdd(p.dstats + (size_t)n * p.stats_nstride + (size_t)k * p.stats_cstride + i, (double)t);
        });
    }
}

// backward pass 1: sums[0]=sum g, [1]=sum g*xhat_a, [2]=sum gb, [3]=sum gb*xhat_b.  The activation derivative is
// recomputed from the pre-activation u = A(a)+actb(B(b)) (same fp32 arithmetic as the forward), so the saved
// output y is never re-read.
template <typename T, bool HAS_B, int ACT, int ACTB>
__global__ void __launch_bounds__(256, HAS_B ? 2 : 3) affine_act_bwd_reduce_kernel(const AffArgs p) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float sm[];  // [rows][4][C] + [<=256]
    constexpr int U = HAS_B ? 2 : SEMB_AFF_U_BWD;
    const int act = ACT >= 0 ? ACT : p.act, actb = ACTB >= 0 ? ACTB : p.actb;
    const Lanes L(p.C);
    const int n = blockIdx.y;
    const int c = L.cg * 8;
    int cb = c;
    const View bv = b_view(p, c, cb, false);
    const long long pix0 = (long long)n * p.HW;
    const int begin = blockIdx.x * p.ppb, end = min(p.HW, begin + p.ppb);
    const size_t aoff = (size_t)n * p.aff_nstride + c;
    const bool red_a = p.mode_a == SEMB_AFF_BATCH, red_b = HAS_B && p.mode_b == SEMB_AFF_BATCH;

    float sa[8], ta[8], ma[8], sb[8], tb[8], mb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { sa[i] = 1.f; ta[i] = 0.f; ma[i] = 0.f; sb[i] = 1.f; tb[i] = 0.f; mb[i] = 0.f; }
    if (L.active) {
        if (p.mode_a != SEMB_AFF_NONE) { ld_params<0>(p.scale_a, aoff, sa); ld_params<0>(p.shift_a, aoff, ta); }
        if (red_a) ld_params<0>(p.mean_a, aoff, ma);
        if (HAS_B && p.mode_b != SEMB_AFF_NONE) { ld_params<0>(p.scale_b, aoff, sb); ld_params<0>(p.shift_b, aoff, tb); }
        if (red_b) ld_params<0>(p.mean_b, aoff, mb);
    }
    float q0[8], q1[8], q2[8], q3[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { q0[i] = q1[i] = q2[i] = q3[i] = 0.f; }

    bool ringed = false;
    if constexpr (sizeof(T) == 2) {
        if (p.ring && L.active) {
            ringed = true;
            const int S = p.ring; constexpr int NOPS = HAS_B ? 3 : 2;
            const uint32_t ring = smem_addr(sm) + (uint32_t)p.ring_off + threadIdx.x * 16u;
            const int first = begin + L.prow;
            const int niter = first < end ? (end - first + L.rows - 1) / L.rows : 0;
            auto issue = [&](int it) {
                if (it < niter) {
                    const long long px = pix0 + first + (long long)it * L.rows;
                    const uint32_t slot = ring + (uint32_t)((it & (S - 1)) * NOPS) * 4096u;
                    cp_async16(slot, vptr<T>(p.dy, px, c));
                    cp_async16(slot + 4096u, vptr<T>(p.a, px, c));
                    if (HAS_B) cp_async16(slot + 8192u, vptr<T>(bv, px, cb));
                }
                cp_async_commit();
            };
            for (int i = 0; i < S - 1; ++i) issue(i);
            for (int it = 0; it < niter; ++it) {
                issue(it + S - 1);
                cp_async_wait_stages(S);
                const uint32_t slot = ring + (uint32_t)((it & (S - 1)) * NOPS) * 4096u;
                Raw8<T> rg, ra, rb;
                rg.r = lds128(slot);
                ra.r = lds128(slot + 4096u);
                if (HAS_B) rb.r = lds128(slot + 8192u);
                float g[8], va[8], vb[8];
                rg.unpack(g);
                ra.unpack(va);
                if (HAS_B) rb.unpack(vb);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float ub = 0.f, t = fmaf(va[i], sa[i], ta[i]);
                    if (HAS_B) { ub = fmaf(vb[i], sb[i], tb[i]); t += act_fwd(ub, actb); }
                    const float gg = g[i] * act_bwd_from_u(t, act);
                    q0[i] += gg;
                    q1[i] += gg * (va[i] - ma[i]);
                    if (HAS_B) {
                        const float gb = gg * act_bwd_from_u(ub, actb);
                        q2[i] += gb;
                        q3[i] += gb * (vb[i] - mb[i]);
                    }
                }
            }
        }
    }
    if (L.active && !ringed) {
        for (int base = begin + L.prow; base < end; base += L.rows * U) {
            Raw8<T> rg[U], ra[U], rb[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int px = base + u * L.rows;
                if (px < end) {
                    rg[u].load(vptr<T>(p.dy, pix0 + px, c));
                    ra[u].load(vptr<T>(p.a, pix0 + px, c));
                    if (HAS_B) rb[u].load(vptr<T>(bv, pix0 + px, cb));
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int px = base + u * L.rows;
                if (px < end) {
                    float g[8], va[8], vb[8];
                    rg[u].unpack(g);
                    ra[u].unpack(va);
                    if (HAS_B) rb[u].unpack(vb);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float ub = 0.f, t = fmaf(va[i], sa[i], ta[i]);
                        if (HAS_B) { ub = fmaf(vb[i], sb[i], tb[i]); t += act_fwd(ub, actb); }
                        const float gg = g[i] * act_bwd_from_u(t, act);
                        q0[i] += gg;
                        q1[i] += gg * (va[i] - ma[i]);
                        if (HAS_B) {
                            const float gb = gg * act_bwd_from_u(ub, actb);
                            q2[i] += gb;
                            q3[i] += gb * (vb[i] - mb[i]);
                        }
                    }
                }
            }
        }
    }
    // block combine without shared-memory atomics: partials [rows][4][C], then fixed-order column sums by all threads
    constexpr int K = HAS_B ? 4 : 2;
    if (L.active) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            sm[(L.prow * K + 0) * p.C + c + i] = q0[i];
            sm[(L.prow * K + 1) * p.C + c + i] = q1[i];
            if (HAS_B) {
                sm[(L.prow * K + 2) * p.C + c + i] = q2[i];
                sm[(L.prow * K + 3) * p.C + c + i] = q3[i];
            }
        }
    }
    float* sm2 = sm + L.rows * K * p.C;
    block_colsum(sm, sm2, L.rows, K * p.C, [&](int col, float t) {
        const int k = col / p.C, i = col - k * p.C;
        float* st = p.stats + (size_t)n * p.stats_nstride + i;
        const size_t ao = (size_t)n * p.aff_nstride + i;
        if (k == 0 && red_a) atomicAdd(st, t);
        if (k == 1 && red_a) atomicAdd(st + p.stats_cstride, t * p.invstd_a[ao]);
        if (k == 2 && red_b) atomicAdd(st + 2 * p.stats_cstride, t);
        if (k == 3 && red_b) atomicAdd(st + 3 * p.stats_cstride, t * p.invstd_b[ao]);
    });
}

// backward pass 2:  da = sa*g + Pa*a + Qa  with  Pa = -sa*inv*c2, Qa = -sa*c1 + sa*inv*c2*mean  (batch-stat norm),
// Pa = Qa = 0 for a constant affine; same for b with gb = g*actb'(ub).
template <typename T, bool HAS_B, int ACT, int ACTB>
__global__ void __launch_bounds__(256, HAS_B ? 2 : 3) affine_act_bwd_apply_kernel(const AffArgs p) {
    pdl_trigger();
    pdl_wait();
    constexpr int U = HAS_B ? 2 : SEMB_AFF_U_BWD;
    const int act = ACT >= 0 ? ACT : p.act, actb = ACTB >= 0 ? ACTB : p.actb;
    const Lanes L(p.C);
    if (!L.active) return;
    const int n = blockIdx.y;
    const int c = L.cg * 8;
    int cb = c;
    const View bv = b_view(p, c, cb, false);
    const long long pix0 = (long long)n * p.HW;
    const int begin = blockIdx.x * p.ppb, end = min(p.HW, begin + p.ppb);
    const size_t aoff = (size_t)n * p.aff_nstride + c;

    float sa[8], ta[8], Pa[8], Qa[8], sb[8], tb[8], Pb[8], Qb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { sa[i] = 1.f; ta[i] = 0.f; Pa[i] = 0.f; Qa[i] = 0.f; sb[i] = 1.f; tb[i] = 0.f; Pb[i] = 0.f; Qb[i] = 0.f; }
    // c1 = sum(g)/count, c2 = sum(g*xhat)/count: from the c1/c2 arrays, or (p.stats != NULL, semb_affine_act_bwd_apply_sums)
    // straight from the sums of pass 1, which folds semb_norm_bwd_finalize into this kernel
    const float* st = p.stats ? p.stats + (size_t)n * p.stats_nstride + c : nullptr;
    if (p.mode_a != SEMB_AFF_NONE) { ld_params<0>(p.scale_a, aoff, sa); ld_params<0>(p.shift_a, aoff, ta); }
    if (p.mode_a == SEMB_AFF_BATCH) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float c1 = st ? st[i] / p.count_a : p.c1_a[aoff + i];
            const float c2 = st ? st[p.stats_cstride + i] / p.count_a : p.c2_a[aoff + i];
            const float k = sa[i] * p.invstd_a[aoff + i] * c2;
            Pa[i] = -k;
            Qa[i] = fmaf(k, p.mean_a[aoff + i], -sa[i] * c1);
        }
    }
    if (HAS_B && p.mode_b != SEMB_AFF_NONE) { ld_params<0>(p.scale_b, aoff, sb); ld_params<0>(p.shift_b, aoff, tb); }
    if (HAS_B && p.mode_b == SEMB_AFF_BATCH) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float c1 = st ? st[2 * p.stats_cstride + i] / p.count_b : p.c1_b[aoff + i];
            const float c2 = st ? st[3 * p.stats_cstride + i] / p.count_b : p.c2_b[aoff + i];
            const float k = sb[i] * p.invstd_b[aoff + i] * c2;
            Pb[i] = -k;
            Qb[i] = fmaf(k, p.mean_b[aoff + i], -sb[i] * c1);
        }
    }
    if (st && blockIdx.x == 0 && L.prow == 0) {      // dgamma += sum(g*xhat), dbeta += sum(g): once per group
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (p.mode_a == SEMB_AFF_BATCH) {
                if (p.dbeta_a) atomicAdd(p.dbeta_a + c + i, st[i]);
                if (p.dgamma_a) atomicAdd(p.dgamma_a + c + i, st[p.stats_cstride + i]);
            }
            if (HAS_B && p.mode_b == SEMB_AFF_BATCH) {
                if (p.dbeta_b) atomicAdd(p.dbeta_b + c + i, st[2 * p.stats_cstride + i]);
                if (p.dgamma_b) atomicAdd(p.dgamma_b + c + i, st[3 * p.stats_cstride + i]);
            }
        }
    }
    int cdb = c;
    const View dbv = b_view(p, c, cdb, true);
    const bool wa = p.da.ptr != nullptr, wb = HAS_B && dbv.ptr != nullptr;
    if (!wa && !wb) return;

    if constexpr (sizeof(T) == 2) {
        if (p.ring) {
            extern __shared__ float sm[];
            const int S = p.ring; constexpr int NOPS = HAS_B ? 3 : 2;
            const uint32_t ring = smem_addr(sm) + (uint32_t)p.ring_off + threadIdx.x * 16u;
            const int first = begin + L.prow;
            const int niter = first < end ? (end - first + L.rows - 1) / L.rows : 0;
            // The reduction pass (same grid, same ranges) walked forwards: the tail of every block's range is what L2 still
            // holds, so this pass walks BACKWARDS and ends on the heads, where the next layer's reduction pass starts.
            const bool rev = p.rev != 0;
            auto issue = [&](int it) {
                if (it < niter) {
                    const long long px = pix0 + first + (long long)(rev ? niter - 1 - it : it) * L.rows;
                    const uint32_t slot = ring + (uint32_t)((it & (S - 1)) * NOPS) * 4096u;
                    cp_async16(slot, vptr<T>(p.dy, px, c));
                    cp_async16(slot + 4096u, vptr<T>(p.a, px, c));
                    if (HAS_B) cp_async16(slot + 8192u, vptr<T>(bv, px, cb));
                }
                cp_async_commit();
            };
            for (int i = 0; i < S - 1; ++i) issue(i);
            for (int it = 0; it < niter; ++it) {
                issue(it + S - 1);
                const long long px = pix0 + first + (long long)(rev ? niter - 1 - it : it) * L.rows;
                // gradients that accumulate (rare: a tensor with a second consumer earlier in the backward order) are read
                // directly, before the wait, so that their latency overlaps the ring's
                float olda[8], oldb[8];
                if (wa && p.acc_a) Vec8<T>::load(vptr<T>(p.da, px, c), olda);
                if (wb && p.acc_b) Vec8<T>::load(vptr<T>(dbv, px, cdb), oldb);
                cp_async_wait_stages(S);
                const uint32_t slot = ring + (uint32_t)((it & (S - 1)) * NOPS) * 4096u;
                Raw8<T> rg, ra, rb;
                rg.r = lds128(slot);
                ra.r = lds128(slot + 4096u);
                if (HAS_B) rb.r = lds128(slot + 8192u);
                float g[8], va[8], vb[8], da[8], db[8];
                rg.unpack(g);
                ra.unpack(va);
                if (HAS_B) rb.unpack(vb);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float ub = 0.f, t = fmaf(va[i], sa[i], ta[i]);
                    if (HAS_B) { ub = fmaf(vb[i], sb[i], tb[i]); t += act_fwd(ub, actb); }
                    const float gg = g[i] * act_bwd_from_u(t, act);
                    da[i] = fmaf(sa[i], gg, fmaf(Pa[i], va[i], Qa[i]));
                    if (HAS_B) {
                        const float gb = gg * act_bwd_from_u(ub, actb);
                        db[i] = fmaf(sb[i], gb, fmaf(Pb[i], vb[i], Qb[i]));
                    }
                }
                if (wa) {
                    if (p.acc_a) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) da[i] += olda[i];
                    }
                    Vec8<T>::store(vptr_mut<T>(p.da, px, c), da);
                }
                if (wb) {
                    if (p.acc_b) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) db[i] += oldb[i];
                    }
                    Vec8<T>::store(vptr_mut<T>(dbv, px, cdb), db);
                }
            }
            return;
        }
    }
    for (int base = begin + L.prow; base < end; base += L.rows * U) {
        Raw8<T> rg[U], ra[U], rb[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int px = base + u * L.rows;
            if (px < end) {
                rg[u].load(vptr<T>(p.dy, pix0 + px, c));
                ra[u].load(vptr<T>(p.a, pix0 + px, c));
                if (HAS_B) rb[u].load(vptr<T>(bv, pix0 + px, cb));
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int px = base + u * L.rows;
            if (px < end) {
                float g[8], va[8], vb[8], da[8], db[8];
                rg[u].unpack(g);
                ra[u].unpack(va);
                if (HAS_B) rb[u].unpack(vb);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float ub = 0.f, t = fmaf(va[i], sa[i], ta[i]);
                    if (HAS_B) { ub = fmaf(vb[i], sb[i], tb[i]); t += act_fwd(ub, actb); }
                    const float gg = g[i] * act_bwd_from_u(t, act);
                    da[i] = fmaf(sa[i], gg, fmaf(Pa[i], va[i], Qa[i]));
                    if (HAS_B) {
                        const float gb = gg * act_bwd_from_u(ub, actb);
                        db[i] = fmaf(sb[i], gb, fmaf(Pb[i], vb[i], Qb[i]));
                    }
                }
                if (wa) {
                    T* o = vptr_mut<T>(p.da, pix0 + px, c);
                    if (p.acc_a) {
                        float old[8];
                        Vec8<T>::load(o, old);
#pragma unroll
                        for (int i = 0; i < 8; ++i) da[i] += old[i];
                    }
                    Vec8<T>::store(o, da);
                }
                if (wb) {
                    T* o = vptr_mut<T>(dbv, pix0 + px, cdb);
                    if (p.acc_b) {
                        float old[8];
                        Vec8<T>::load(o, old);
#pragma unroll
                        for (int i = 0; i < 8; ++i) db[i] += old[i];
                    }
                    Vec8<T>::store(o, db);
                }
            }
        }
    }
}

// Fused backward of the normalisation + activation: pass 1 (sums), a grid-wide barrier, pass 2 (gradients) in ONE
// cooperative launch.  Pass 2 walks each block's pixel range BACKWARDS, so that the part of dy / a / b the block read
// last in pass 1 is re-read while it is still in the 126 MB L2 -- for the layers whose operands fit L2 the second pass
// costs no HBM reads at all; the C-length finalize (c1, c2, dgamma, dbeta) is folded in as well (three launches -> one).
template <typename T, bool HAS_B, int ACT, int ACTB>
__global__ void __launch_bounds__(256, HAS_B ? 2 : 3) affine_act_bwd_fused_kernel(const AffArgs p) {
    extern __shared__ float sm[];  // [rows][K][C] + [<=256]
    constexpr int U = 2;
    const int act = ACT >= 0 ? ACT : p.act, actb = ACTB >= 0 ? ACTB : p.actb;
    const Lanes L(p.C);
    const int n = blockIdx.y;
    const int c = L.cg * 8;
    int cb = c;
    const View bv = b_view(p, c, cb, false);
    const long long pix0 = (long long)n * p.HW;
    const int begin = blockIdx.x * p.ppb, end = min(p.HW, begin + p.ppb);
    const size_t aoff = (size_t)n * p.aff_nstride + c;
    const bool red_a = p.mode_a == SEMB_AFF_BATCH, red_b = HAS_B && p.mode_b == SEMB_AFF_BATCH;

    float sa[8], ta[8], ma[8], sb[8], tb[8], mb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { sa[i] = 1.f; ta[i] = 0.f; ma[i] = 0.f; sb[i] = 1.f; tb[i] = 0.f; mb[i] = 0.f; }
    if (L.active) {
        if (p.mode_a != SEMB_AFF_NONE) { ld_params<0>(p.scale_a, aoff, sa); ld_params<0>(p.shift_a, aoff, ta); }
        if (red_a) ld_params<0>(p.mean_a, aoff, ma);
        if (HAS_B && p.mode_b != SEMB_AFF_NONE) { ld_params<0>(p.scale_b, aoff, sb); ld_params<0>(p.shift_b, aoff, tb); }
        if (red_b) ld_params<0>(p.mean_b, aoff, mb);
    }
    // ---------------- pass 1: sums ----------------
    {
        float q0[8], q1[8], q2[8], q3[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { q0[i] = q1[i] = q2[i] = q3[i] = 0.f; }
        if (L.active) {
            for (int base = begin + L.prow; base < end; base += L.rows * U) {
                Raw8<T> rg[U], ra[U], rb[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int px = base + u * L.rows;
                    if (px < end) {
                        rg[u].load(vptr<T>(p.dy, pix0 + px, c));
                        ra[u].load(vptr<T>(p.a, pix0 + px, c));
                        if (HAS_B) rb[u].load(vptr<T>(bv, pix0 + px, cb));
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int px = base + u * L.rows;
                    if (px < end) {
                        float g[8], va[8], vb[8];
                        rg[u].unpack(g);
                        ra[u].unpack(va);
                        if (HAS_B) rb[u].unpack(vb);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float ub = 0.f, t = fmaf(va[i], sa[i], ta[i]);
                            if (HAS_B) { ub = fmaf(vb[i], sb[i], tb[i]); t += act_fwd(ub, actb); }
                            const float gg = g[i] * act_bwd_from_u(t, act);
                            q0[i] += gg;
                            q1[i] += gg * (va[i] - ma[i]);
                            if (HAS_B) {
                                const float gb = gg * act_bwd_from_u(ub, actb);
                                q2[i] += gb;
                                q3[i] += gb * (vb[i] - mb[i]);
                            }
                        }
                    }
                }
            }
        }
        constexpr int K = HAS_B ? 4 : 2;
        if (L.active) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                sm[(L.prow * K + 0) * p.C + c + i] = q0[i];
                sm[(L.prow * K + 1) * p.C + c + i] = q1[i];
                if (HAS_B) {
                    sm[(L.prow * K + 2) * p.C + c + i] = q2[i];
                    sm[(L.prow * K + 3) * p.C + c + i] = q3[i];
                }
            }
        }
        float* sm2 = sm + L.rows * K * p.C;
        block_colsum(sm, sm2, L.rows, K * p.C, [&](int col, float t) {
            const int k = col / p.C, i = col - k * p.C;
            float* st = p.stats + (size_t)n * p.stats_nstride + i;
            const size_t ao = (size_t)n * p.aff_nstride + i;
            if (k == 0 && red_a) atomicAdd(st, t);
            if (k == 1 && red_a) atomicAdd(st + p.stats_cstride, t * p.invstd_a[ao]);
            if (k == 2 && red_b) atomicAdd(st + 2 * p.stats_cstride, t);
            if (k == 3 && red_b) atomicAdd(st + 3 * p.stats_cstride, t * p.invstd_b[ao]);
        });
    }
    // ---------------- grid-wide barrier (cooperative launch: all blocks are co-resident) ----------------
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int expected = gridDim.x * gridDim.y;
        atomicAdd(p.barrier, 1u);
        while (*reinterpret_cast<volatile unsigned int*>(p.barrier) < expected) __nanosleep(64);
        __threadfence();
        // self-resetting: the last block to LEAVE the barrier clears both words for the next launch
        if (atomicAdd(p.barrier + 1, 1u) == expected - 1) { p.barrier[0] = 0u; p.barrier[1] = 0u; }
    }
    __syncthreads();
    // ---------------- C-length finalize: dgamma / dbeta (one block per sample row) ----------------
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < p.C; i += 256) {
            const float* st = p.stats + (size_t)n * p.stats_nstride + i;
            if (red_a) {
                if (p.dbeta_a) atomicAdd(p.dbeta_a + i, __ldcg(st));
                if (p.dgamma_a) atomicAdd(p.dgamma_a + i, __ldcg(st + p.stats_cstride));
            }
            if (red_b) {
                if (p.dbeta_b) atomicAdd(p.dbeta_b + i, __ldcg(st + 2 * p.stats_cstride));
                if (p.dgamma_b) atomicAdd(p.dgamma_b + i, __ldcg(st + 3 * p.stats_cstride));
            }
        }
    }
    if (!L.active) return;
    // ---------------- pass 2: gradients, newest data first ----------------
    float Pa[8], Qa[8], Pb[8], Qb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { Pa[i] = 0.f; Qa[i] = 0.f; Pb[i] = 0.f; Qb[i] = 0.f; }
    {
        const float* st = p.stats + (size_t)n * p.stats_nstride + c;
        if (red_a) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float c1 = __ldcg(st + i) / p.count_a, c2 = __ldcg(st + p.stats_cstride + i) / p.count_a;
                const float k = sa[i] * p.invstd_a[aoff + i] * c2;
                Pa[i] = -k;
                Qa[i] = fmaf(k, ma[i], -sa[i] * c1);
            }
        }
        if (red_b) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float c1 = __ldcg(st + 2 * p.stats_cstride + i) / p.count_b, c2 = __ldcg(st + 3 * p.stats_cstride + i) / p.count_b;
                const float k = sb[i] * p.invstd_b[aoff + i] * c2;
                Pb[i] = -k;
                Qb[i] = fmaf(k, mb[i], -sb[i] * c1);
            }
        }
    }
    int cdb = c;
    const View dbv = b_view(p, c, cdb, true);
    const bool wa = p.da.ptr != nullptr, wb = HAS_B && dbv.ptr != nullptr;
    if (!wa && !wb) return;
    const int span = L.rows * U;
    const int niter = (end - begin - L.prow + span - 1) / span;      // iterations of this thread in pass 1
    for (int it = niter - 1; it >= 0; --it) {
        const int base = begin + L.prow + it * span;
        Raw8<T> rg[U], ra[U], rb[U];
#pragma unroll
        for (int u = U - 1; u >= 0; --u) {
            const int px = base + u * L.rows;
            if (px < end) {
                rg[u].load(vptr<T>(p.dy, pix0 + px, c));
                ra[u].load(vptr<T>(p.a, pix0 + px, c));
                if (HAS_B) rb[u].load(vptr<T>(bv, pix0 + px, cb));
            }
        }
#pragma unroll
        for (int u = U - 1; u >= 0; --u) {
            const int px = base + u * L.rows;
            if (px < end) {
                float g[8], va[8], vb[8], da[8], db[8];
                rg[u].unpack(g);
                ra[u].unpack(va);
                if (HAS_B) rb[u].unpack(vb);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float ub = 0.f, t = fmaf(va[i], sa[i], ta[i]);
                    if (HAS_B) { ub = fmaf(vb[i], sb[i], tb[i]); t += act_fwd(ub, actb); }
                    const float gg = g[i] * act_bwd_from_u(t, act);
                    da[i] = fmaf(sa[i], gg, fmaf(Pa[i], va[i], Qa[i]));
                    if (HAS_B) {
                        const float gb = gg * act_bwd_from_u(ub, actb);
                        db[i] = fmaf(sb[i], gb, fmaf(Pb[i], vb[i], Qb[i]));
                    }
                }
                if (wa) {
                    T* o = vptr_mut<T>(p.da, pix0 + px, c);
                    if (p.acc_a) {
                        float old[8];
                        Vec8<T>::load(o, old);
#pragma unroll
                        for (int i = 0; i < 8; ++i) da[i] += old[i];
                    }
                    Vec8<T>::store(o, da);
                }
                if (wb) {
                    T* o = vptr_mut<T>(dbv, pix0 + px, cdb);
                    if (p.acc_b) {
                        float old[8];
                        Vec8<T>::load(o, old);
#pragma unroll
                        for (int i = 0; i < 8; ++i) db[i] += old[i];
                    }
                    Vec8<T>::store(o, db);
                }
            }
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) channel_sum_kernel(const AffArgs p) {
    extern __shared__ float sm[];  // [C]
    const Lanes L(p.C);
    const int n = blockIdx.y;
    const int c = L.cg * 8;
    const long long pix0 = (long long)n * p.HW;
    const int begin = blockIdx.x * p.ppb, end = min(p.HW, begin + p.ppb);
    float s1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s1[i] = 0.f;
    if (L.active) {
        for (int px = begin + L.prow; px < end; px += L.rows) {
            float va[8];
            Vec8<T>::load(vptr<T>(p.a, pix0 + px, c), va);
#pragma unroll
            for (int i = 0; i < 8; ++i) s1[i] += va[i];
        }
    }
    for (int i = threadIdx.x; i < p.C; i += 256) sm[i] = 0.f;
    __syncthreads();
    if (L.active) {
#pragma unroll
        for (int i = 0; i < 8; ++i) atomicAdd(&sm[c + i], s1[i]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < p.C; i += 256) atomicAdd(p.stats + i, sm[i]);
}

static int pick_ppb(int HW, int N, int C) {
    const int rows = 256 / (C / 8);
    // Per-sample (InstanceNorm) grids are cut into ~148 * 3 blocks of at least 8 block iterations each.  Every block derives the
    // scale / shift of its channels from the fp64 moments before its first pixel (the folded finalize), so short blocks are
    // prologue-bound: measured on the CycleGAN step [B200] 148 * 12 blocks x >= 4 iterations 10.6 ms of affine kernels per step,
    // 148 * 3 x >= 8: 8.0 ms (step 52.6 -> 50.1 ms); 148 * 1 x >= 32: 11.0 ms.  SEMB_AFF_PS="blocks_per_sm,min_iterations" overrides.
    static const int knob_blocks = [] { const char* e = getenv("SEMB_AFF_PS"); int a = 3, b = 8; if (e) sscanf(e, "%d,%d", &a, &b); return a > 0 ? a : 3; }();
    static const int knob_iters = [] { const char* e = getenv("SEMB_AFF_PS"); int a = 3, b = 8; if (e) sscanf(e, "%d,%d", &a, &b); return b > 0 ? b : 8; }();
    long long target_blocks = 148LL * knob_blocks;
    int ppb = (int)cdivl((long long)HW * N, target_blocks);
    int min_ppb = rows * knob_iters;
    if (ppb < min_ppb) ppb = min_ppb;
    ppb = cdiv(ppb, rows) * rows;
    return ppb;
}

// Grid of the affine kernels.  With per-channel (BatchNorm) parameters the batch is one flat pixel range, cut into
// exactly ONE wave of blocks (148 SMs x resident blocks): each block then amortises its block-level reduction and
// atomics over many pixels (round-1 profile: the 8/16-channel full-resolution layers ran at 2 TB/s with 1280-pixel
// blocks).  Per-sample (InstanceNorm) parameters keep one grid row per sample.
static dim3 aff_grid(AffArgs& p, bool per_sample, int blocks_per_sm) {
    const int rows = 256 / (p.C / 8);
    if (!per_sample) {
        const long long total = (long long)p.N * p.HW;
        if (total < (1LL << 31)) {
            p.N = 1;
            p.HW = (int)total;
            // at least 8 block iterations per block (same prologue argument as pick_ppb; 4 -> 8: 11.535 -> 11.52 ms per UNet step)
            static const int bn_iters = [] { const char* e = getenv("SEMB_AFF_BN_ITERS"); const int v = e ? atoi(e) : 8; return v > 0 ? v : 8; }();
            int ppb = (int)cdivl(total, 148LL * blocks_per_sm);
            if (ppb < rows * bn_iters) ppb = rows * bn_iters;
            p.ppb = cdiv(ppb, rows) * rows;
            return dim3(cdiv(p.HW, p.ppb), 1);
        }
    }
    p.ppb = pick_ppb(p.HW, p.N, p.C);
    return dim3(cdiv(p.HW, p.ppb), p.N);
}

static void set_segments(AffArgs& p, const semb_affine_desc* d) {
    p.nseg_b = d->nseg_b;
    for (int s = 0; s < 3; ++s) {
        p.seg_c0[s] = d->seg_c0[s];
        p.seg_b[s] = View{d->seg_b[s].ptr, d->seg_b[s].pitch, d->seg_b[s].coff};
        p.seg_db[s] = View{d->seg_db[s].ptr, d->seg_db[s].pitch, d->seg_db[s].coff};
    }
}

static int check_aff(const semb_affine_desc* d) {
    SEMB_REQUIRE(d, SEMB_ESHAPE, "affine: null desc");
    SEMB_REQUIRE(d->N > 0 && d->HW > 0 && d->C > 0 && d->C % 8 == 0 && d->C <= 2048, SEMB_ESHAPE,
                 "affine: bad shape N=%d HW=%d C=%d", d->N, d->HW, d->C);
    SEMB_REQUIRE(d->dtype == SEMB_F32 || d->dtype == SEMB_BF16, SEMB_ESHAPE, "affine: bad dtype");
    SEMB_REQUIRE(d->nseg_b >= 0 && d->nseg_b <= 3, SEMB_ESHAPE, "affine: at most three b segments");
    for (int s = 0, c0 = 0; s < d->nseg_b; ++s) {
        SEMB_REQUIRE(view_ok(&d->seg_b[s]) && d->seg_c0[s] == c0, SEMB_ESHAPE, "affine: b segments must be valid views that tile the channels in order");
        c0 += d->seg_b[s].C;
        SEMB_REQUIRE(s + 1 < d->nseg_b || c0 == d->C, SEMB_ESHAPE, "affine: b segments cover %d of %d channels", c0, d->C);
        SEMB_REQUIRE(!d->seg_db[s].ptr || (view_ok(&d->seg_db[s]) && d->seg_db[s].C == d->seg_b[s].C), SEMB_ESHAPE, "affine: bad b gradient segment");
    }
    return SEMB_OK;
}

__global__ void norm_finalize_kernel(const double* stats, int groups, int C, int cstride, int stats_nstride, float count,
                                     float eps, const float* gamma, const float* beta, float* scale, float* shift,
                                     float* mean_o, float* invstd_o, float* mm, float* mv, float momentum) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= groups * C) return;
    const int g = i / C, c = i % C;
    const double* st = stats + (size_t)g * stats_nstride;
    const double dm = st[c] / (double)count;
    const float mean = (float)dm;
    const float var = (float)(st[cstride + c] / (double)count - dm * dm);   // Keras: E[x^2] - E[x]^2 (biased)
    const float inv = rsqrtf(var + eps);
    const float ga = gamma ? gamma[c] : 1.f;
    const float sc = ga * inv;
    const size_t o = (size_t)g * cstride + c;
    scale[o] = sc;
    shift[o] = beta[c] - mean * sc;
    if (mean_o) mean_o[o] = mean;
    if (invstd_o) invstd_o[o] = inv;
    if (mm) {
        mm[c] = mm[c] * momentum + mean * (1.f - momentum);
        mv[c] = mv[c] * momentum + var * (1.f - momentum);
    }
}

__global__ void norm_from_moving_kernel(int C, float eps, const float* gamma, const float* beta, const float* mm,
                                        const float* mv, float* scale, float* shift) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float sc = (gamma ? gamma[c] : 1.f) * rsqrtf(mv[c] + eps);
    scale[c] = sc;
    shift[c] = beta[c] - mm[c] * sc;
}

__global__ void norm_bwd_finalize_kernel(const float* sums, int which, int groups, int C, int cstride, int sums_nstride,
                                         float count, float* c1, float* c2, float* dgamma, float* dbeta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float tg = 0.f, tb = 0.f;
    for (int g = 0; g < groups; ++g) {
        const float* s = sums + (size_t)g * sums_nstride + (size_t)(2 * which) * cstride;
        const float s0 = s[c], s1 = s[cstride + c];
        c1[(size_t)g * cstride + c] = s0 / count;
        c2[(size_t)g * cstride + c] = s1 / count;
        tb += s0;
        tg += s1;
    }
    if (dgamma) dgamma[c] += tg;
    if (dbeta) dbeta[c] += tb;
}

}  // namespace semb

using namespace semb;

extern "C" int semb_norm_finalize(const void* stats, int32_t groups, int32_t C, int32_t cstride, int32_t stats_nstride,
                                  float count, float eps, const float* gamma, const float* beta, float* scale,
                                  float* shift, float* mean, float* invstd, float* moving_mean, float* moving_var,
                                  float momentum, void* stream) {
    SEMB_REQUIRE(stats && beta && scale && shift && groups > 0 && C > 0, SEMB_ESHAPE, "norm_finalize: bad arguments");
    SEMB_REQUIRE(!moving_mean || groups == 1, SEMB_ESHAPE, "norm_finalize: moving statistics need groups==1");
    const int n = groups * C;
    norm_finalize_kernel<<<cdiv(n, 128), 128, 0, as_stream(stream)>>>(reinterpret_cast<const double*>(stats), groups, C, cstride, stats_nstride, count, eps,
                                                                     gamma, beta, scale, shift, mean, invstd,
                                                                     moving_mean, moving_var, momentum);
    return check_launch("norm_finalize");
}

extern "C" int semb_norm_from_moving(int32_t C, float eps, const float* gamma, const float* beta, const float* moving_mean,
                                     const float* moving_var, float* scale, float* shift, void* stream) {
    SEMB_REQUIRE(C > 0 && beta && moving_mean && moving_var && scale && shift, SEMB_ESHAPE, "norm_from_moving: bad arguments");
    norm_from_moving_kernel<<<cdiv(C, 128), 128, 0, as_stream(stream)>>>(C, eps, gamma, beta, moving_mean, moving_var, scale, shift);
    return check_launch("norm_from_moving");
}

extern "C" int semb_norm_bwd_finalize(const float* sums, int32_t which, int32_t groups, int32_t C, int32_t cstride,
                                      int32_t sums_nstride, float count, const float* mean, const float* invstd,
                                      float* c1, float* c2, float* dgamma, float* dbeta, void* stream) {
    (void)mean; (void)invstd;
    SEMB_REQUIRE(sums && c1 && c2 && groups > 0 && C > 0 && (which == 0 || which == 1), SEMB_ESHAPE, "norm_bwd_finalize: bad arguments");
    norm_bwd_finalize_kernel<<<cdiv(C, 128), 128, 0, as_stream(stream)>>>(sums, which, groups, C, cstride, sums_nstride,
                                                                         count, c1, c2, dgamma, dbeta);
    return check_launch("norm_bwd_finalize");
}


// every affine kernel starts with pdl_trigger() / pdl_wait() (the fused cooperative one is launched elsewhere)
#define SEMB_AFF_GO(grid, smem, st, p, ...)                                                                         \
    do {                                                                                                            \
        if ((smem) > 48 * 1024) cudaFuncSetAttribute(__VA_ARGS__, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem)); \
        launch_pdl(__VA_ARGS__, dim3(grid), dim3(256), smem, st, p);                                                \
    } while (0)

// Operand ring of the bf16 kernels (OperandRing above): appends `stages * nops` 4 KB slot planes to the kernel's dynamic
// shared memory and returns the total; `blocks` is lowered to what then fits an SM.  SEMB_AFF_NO_RING=1: register prefetch.
enum { RING_FWD = 0, RING_FWDB, RING_RED, RING_REDB, RING_APP, RING_APPB };
static int ring_stages(int which) {
    // defaults from the B200 measurements above; SEMB_AFF_RING="fwd,fwdb,red,redb,app,appb" overrides (0 / 2 / 4 / 8 each)
    static int st[6] = {8, 8, 4, 4, 8, 8};
    static const bool init = [] {
        if (const char* e = getenv("SEMB_AFF_RING")) {
            int v[6];
            if (sscanf(e, "%d,%d,%d,%d,%d,%d", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5]) == 6)
                for (int i = 0; i < 6; ++i) if (v[i] == 0 || v[i] == 2 || v[i] == 4 || v[i] == 8) st[i] = v[i];
        }
        if (const char* e = getenv("SEMB_AFF_NO_RING")) if (e[0] && e[0] != '0') for (int i = 0; i < 6; ++i) st[i] = 0;
        return true;
    }();
    (void)init;
    return st[which];
}
static size_t ring_setup(AffArgs& p, int dtype, size_t base_bytes, int stages, int nops, int* blocks) {
    p.ring = 0;
    p.ring_off = 0;
    if (dtype != SEMB_BF16 || stages == 0) return base_bytes;
    const size_t off = (base_bytes + 15) / 16 * 16;
    p.ring = stages;
    p.ring_off = (int)off;
    const size_t total = off + (size_t)stages * nops * 4096;
    const int fit = (int)((224 * 1024) / (total + 1024));
    if (*blocks > fit) *blocks = fit < 1 ? 1 : fit;
    return total;
}

// (activation, second-operand activation) combinations compiled statically; anything else takes the runtime path
#define SEMB_AFF_DISPATCH(KERNEL, T, HASB, grid, smem, st, p)                                                       \
    do {                                                                                                            \
        if ((p).act == SEMB_ACT_NONE && (!(HASB) || (p).actb == SEMB_ACT_NONE))                                     \
            SEMB_AFF_GO(grid, smem, st, p, KERNEL<T, HASB, SEMB_ACT_NONE, SEMB_ACT_NONE>);                              \
        else if ((p).act == SEMB_ACT_RELU && (!(HASB) || (p).actb == SEMB_ACT_NONE))                                \
            SEMB_AFF_GO(grid, smem, st, p, KERNEL<T, HASB, SEMB_ACT_RELU, SEMB_ACT_NONE>);                              \
        else if ((p).act == SEMB_ACT_RELU && (p).actb == SEMB_ACT_RELU)                                             \
            SEMB_AFF_GO(grid, smem, st, p, KERNEL<T, HASB, SEMB_ACT_RELU, SEMB_ACT_RELU>);                              \
        else                                                                                                        \
            SEMB_AFF_GO(grid, smem, st, p, KERNEL<T, HASB, -1, -1>);                                                    \
    } while (0)

#define SEMB_AFF_LAUNCH(KERNEL, dtype, has_b, grid, smem, st, p)                                                    \
    do {                                                                                                            \
        if ((dtype) == SEMB_BF16) {                                                                                 \
            if (has_b) SEMB_AFF_DISPATCH(KERNEL, bf16, true, grid, smem, st, p);                                    \
            else SEMB_AFF_DISPATCH(KERNEL, bf16, false, grid, smem, st, p);                                         \
        } else {                                                                                                    \
            if (has_b) SEMB_AFF_DISPATCH(KERNEL, float, true, grid, smem, st, p);                                   \
            else SEMB_AFF_DISPATCH(KERNEL, float, false, grid, smem, st, p);                                        \
        }                                                                                                           \
    } while (0)

extern "C" int semb_affine_act_fwd(const semb_affine_desc* d, const semb_tensor* a, const float* scale_a,
                                   const float* shift_a, const semb_tensor* b, const float* scale_b, const float* shift_b,
                                   const semb_tensor* y, void* stats, int32_t stats_nstride, int32_t stats_cstride,
                                   void* stream) {
    int rc = check_aff(d);
    if (rc) return rc;
    SEMB_REQUIRE(view_ok(a) && view_ok(y) && (!b || view_ok(b)), SEMB_EALIGN, "affine fwd: bad tensor view");
    SEMB_REQUIRE(a->C == d->C && y->C == d->C && (!b || b->C == d->C), SEMB_ESHAPE, "affine fwd: channel mismatch");
    SEMB_REQUIRE(d->mode_a == SEMB_AFF_NONE || (scale_a && shift_a), SEMB_ESHAPE, "affine fwd: missing scale/shift a");
    SEMB_REQUIRE(!b || d->mode_b == SEMB_AFF_NONE || (scale_b && shift_b), SEMB_ESHAPE, "affine fwd: missing scale/shift b");
    AffArgs p{};
    p.N = d->N; p.HW = d->HW; p.C = d->C; p.act = d->act; p.actb = d->actb; p.mode_a = d->mode_a; p.mode_b = d->mode_b;
    p.aff_nstride = d->aff_nstride;
    set_segments(p, d);
    p.a = mkview(a); p.b = mkview(b); p.y = mkview(y);
    p.scale_a = scale_a; p.shift_a = shift_a; p.scale_b = scale_b; p.shift_b = shift_b;
    p.dstats = reinterpret_cast<double*>(stats); p.stats_nstride = stats_nstride; p.stats_cstride = stats_cstride;
    int blocks = 3;
    const size_t smem = ring_setup(p, d->dtype, stats ? ((size_t)(256 / (d->C / 8)) * 2 * d->C + 256) * sizeof(float) : 0, ring_stages(b ? RING_FWDB : RING_FWD), b ? 2 : 1, &blocks);
    dim3 grid = aff_grid(p, d->aff_nstride != 0 || (stats && stats_nstride != 0), blocks);
    cudaStream_t st = as_stream(stream);
    SEMB_AFF_LAUNCH(affine_act_fwd_kernel, d->dtype, b != nullptr, grid, smem, st, p);
    return check_launch("affine_act_fwd");
}

extern "C" int semb_affine_act_bwd_reduce(const semb_affine_desc* d, const semb_tensor* dy, const semb_tensor* a,
                                          const semb_tensor* b, const float* scale_a, const float* shift_a,
                                          const float* mean_a, const float* invstd_a, const float* scale_b,
                                          const float* shift_b, const float* mean_b, const float* invstd_b, float* sums,
                                          int32_t sums_nstride, int32_t sums_cstride, void* stream) {
    int rc = check_aff(d);
    if (rc) return rc;
    SEMB_REQUIRE(view_ok(dy) && view_ok(a) && (!b || view_ok(b)), SEMB_EALIGN, "affine bwd reduce: bad tensor view");
    SEMB_REQUIRE(sums, SEMB_ESHAPE, "affine bwd reduce: null sums");
    SEMB_REQUIRE(d->mode_a == SEMB_AFF_NONE || (scale_a && shift_a), SEMB_ESHAPE, "affine bwd reduce: missing scale/shift a");
    SEMB_REQUIRE(d->mode_a != SEMB_AFF_BATCH || (mean_a && invstd_a), SEMB_ESHAPE, "affine bwd reduce: missing mean/invstd a");
    SEMB_REQUIRE(!b || d->mode_b == SEMB_AFF_NONE || (scale_b && shift_b), SEMB_ESHAPE, "affine bwd reduce: missing scale/shift b");
    SEMB_REQUIRE(!b || d->mode_b != SEMB_AFF_BATCH || (mean_b && invstd_b), SEMB_ESHAPE, "affine bwd reduce: missing mean/invstd b");
    AffArgs p{};
    p.N = d->N; p.HW = d->HW; p.C = d->C; p.act = d->act; p.actb = d->actb; p.mode_a = d->mode_a; p.mode_b = d->mode_b;
    p.aff_nstride = d->aff_nstride;
    set_segments(p, d);
    p.a = mkview(a); p.b = mkview(b); p.dy = mkview(dy);
    p.scale_a = scale_a; p.shift_a = shift_a; p.mean_a = mean_a; p.invstd_a = invstd_a;
    p.scale_b = scale_b; p.shift_b = shift_b; p.mean_b = mean_b; p.invstd_b = invstd_b;
    p.stats = sums; p.stats_nstride = sums_nstride; p.stats_cstride = sums_cstride;
    int blocks = b ? 2 : 3;
    const size_t smem = ring_setup(p, d->dtype, ((size_t)(256 / (d->C / 8)) * (b ? 4 : 2) * d->C + 256) * sizeof(float), ring_stages(b ? RING_REDB : RING_RED), b ? 3 : 2, &blocks);
    dim3 grid = aff_grid(p, d->aff_nstride != 0 || sums_nstride != 0, blocks);
    cudaStream_t st = as_stream(stream);
    SEMB_AFF_LAUNCH(affine_act_bwd_reduce_kernel, d->dtype, b != nullptr, grid, smem, st, p);
    return check_launch("affine_act_bwd_reduce");
}

extern "C" int semb_affine_act_bwd_apply(const semb_affine_desc* d, const semb_tensor* dy, const semb_tensor* a,
                                         const semb_tensor* b, const float* scale_a, const float* shift_a,
                                         const float* mean_a, const float* invstd_a, const float* c1_a, const float* c2_a,
                                         const float* scale_b, const float* shift_b, const float* mean_b,
                                         const float* invstd_b, const float* c1_b, const float* c2_b,
                                         const semb_tensor* da, int32_t acc_a, const semb_tensor* db, int32_t acc_b,
                                         void* stream) {
    int rc = check_aff(d);
    if (rc) return rc;
    SEMB_REQUIRE(view_ok(dy) && view_ok(a) && (!b || view_ok(b)), SEMB_EALIGN, "affine bwd apply: bad dy/a/b view");
    SEMB_REQUIRE((!da || view_ok(da)) && (!db || view_ok(db)), SEMB_EALIGN, "affine bwd apply: bad gradient view");
    SEMB_REQUIRE(d->mode_a == SEMB_AFF_NONE || (scale_a && shift_a), SEMB_ESHAPE, "affine bwd apply: missing scale/shift a");
    SEMB_REQUIRE(d->mode_a != SEMB_AFF_BATCH || (mean_a && invstd_a && c1_a && c2_a), SEMB_ESHAPE,
                 "affine bwd apply: missing batch-norm terms for a");
    SEMB_REQUIRE(!b || d->mode_b == SEMB_AFF_NONE || (scale_b && shift_b), SEMB_ESHAPE, "affine bwd apply: missing scale/shift b");
    SEMB_REQUIRE(!b || d->mode_b != SEMB_AFF_BATCH || (mean_b && invstd_b && c1_b && c2_b), SEMB_ESHAPE,
                 "affine bwd apply: missing batch-norm terms for b");
    AffArgs p{};
    p.N = d->N; p.HW = d->HW; p.C = d->C; p.act = d->act; p.actb = d->actb; p.mode_a = d->mode_a; p.mode_b = d->mode_b;
    p.aff_nstride = d->aff_nstride;
    set_segments(p, d);
    p.a = mkview(a); p.b = mkview(b); p.dy = mkview(dy); p.da = mkview(da); p.db = mkview(db);
    p.scale_a = scale_a; p.shift_a = shift_a; p.mean_a = mean_a; p.invstd_a = invstd_a; p.c1_a = c1_a; p.c2_a = c2_a;
    p.scale_b = scale_b; p.shift_b = shift_b; p.mean_b = mean_b; p.invstd_b = invstd_b; p.c1_b = c1_b; p.c2_b = c2_b;
    p.acc_a = acc_a; p.acc_b = acc_b;
    int blocks = b ? 2 : 3;
    static const int walk_back = [] { const char* e = getenv("SEMB_AFF_APPLY_FORWARD"); return (e && e[0] && e[0] != '0') ? 0 : 1; }();
    p.rev = walk_back;
    const size_t smem = ring_setup(p, d->dtype, 0, ring_stages(b ? RING_APPB : RING_APP), b ? 3 : 2, &blocks);
    dim3 grid = aff_grid(p, d->aff_nstride != 0, blocks);
    cudaStream_t st = as_stream(stream);
    SEMB_AFF_LAUNCH(affine_act_bwd_apply_kernel, d->dtype, b != nullptr, grid, smem, st, p);
    return check_launch("affine_act_bwd_apply");
}

// ---- forward with the normalisation finalize folded in -----------------------------------------------------------------
extern "C" int semb_affine_act_fwd_fin(const semb_affine_desc* d, const semb_tensor* a, const semb_norm_fin* fin_a,
                                       const semb_tensor* b, const semb_norm_fin* fin_b, const semb_tensor* y, void* stats,
                                       int32_t stats_nstride, int32_t stats_cstride, void* stream) {
    int rc = check_aff(d);
    if (rc) return rc;
    SEMB_REQUIRE(view_ok(a) && view_ok(y) && (!b || view_ok(b)), SEMB_EALIGN, "affine fwd fin: bad tensor view");
    SEMB_REQUIRE(a->C == d->C && y->C == d->C && (!b || b->C == d->C), SEMB_ESHAPE, "affine fwd fin: channel mismatch");
    SEMB_REQUIRE(d->mode_a == SEMB_AFF_NONE || (fin_a && fin_a->scale && fin_a->shift && (!fin_a->stats || fin_a->beta)), SEMB_ESHAPE,
                 "affine fwd fin: missing normalisation record for a");
    SEMB_REQUIRE(!b || d->mode_b == SEMB_AFF_NONE || (fin_b && fin_b->scale && fin_b->shift && (!fin_b->stats || fin_b->beta)), SEMB_ESHAPE,
                 "affine fwd fin: missing normalisation record for b");
    AffArgs p{};
    p.N = d->N; p.HW = d->HW; p.C = d->C; p.act = d->act; p.actb = d->actb; p.mode_a = d->mode_a; p.mode_b = d->mode_b;
    p.aff_nstride = d->aff_nstride;
    set_segments(p, d);
    p.a = mkview(a); p.b = mkview(b); p.y = mkview(y);
    if (fin_a) { p.fa = *fin_a; p.scale_a = fin_a->scale; p.shift_a = fin_a->shift; }
    if (fin_b) { p.fb = *fin_b; p.scale_b = fin_b->scale; p.shift_b = fin_b->shift; }
    p.dstats = reinterpret_cast<double*>(stats); p.stats_nstride = stats_nstride; p.stats_cstride = stats_cstride;
    int blocks = 3;
    const size_t smem = ring_setup(p, d->dtype, stats ? ((size_t)(256 / (d->C / 8)) * 2 * d->C + 256) * sizeof(float) : 0, ring_stages(b ? RING_FWDB : RING_FWD), b ? 2 : 1, &blocks);
    dim3 grid = aff_grid(p, d->aff_nstride != 0 || (stats && stats_nstride != 0), blocks);
    cudaStream_t st = as_stream(stream);
    SEMB_AFF_LAUNCH(affine_act_fwd_kernel, d->dtype, b != nullptr, grid, smem, st, p);
    return check_launch("affine_act_fwd_fin");
}

// ---- backward pass 2 straight from the sums of pass 1 (folds semb_norm_bwd_finalize) ---------------------------------------
extern "C" int semb_affine_act_bwd_apply_sums(const semb_affine_desc* d, const semb_tensor* dy, const semb_tensor* a, const semb_tensor* b,
                                              const float* scale_a, const float* shift_a, const float* mean_a, const float* invstd_a,
                                              float count_a, float* dgamma_a, float* dbeta_a,
                                              const float* scale_b, const float* shift_b, const float* mean_b, const float* invstd_b,
                                              float count_b, float* dgamma_b, float* dbeta_b,
                                              const float* sums, int32_t sums_nstride, int32_t sums_cstride,
                                              const semb_tensor* da, int32_t acc_a, const semb_tensor* db, int32_t acc_b, void* stream) {
    int rc = check_aff(d);
    if (rc) return rc;
    SEMB_REQUIRE(view_ok(dy) && view_ok(a) && (!b || view_ok(b)), SEMB_EALIGN, "affine bwd apply sums: bad dy/a/b view");
    SEMB_REQUIRE((!da || view_ok(da)) && (!db || view_ok(db)), SEMB_EALIGN, "affine bwd apply sums: bad gradient view");
    SEMB_REQUIRE(sums, SEMB_ESHAPE, "affine bwd apply sums: null sums");
    SEMB_REQUIRE(d->mode_a == SEMB_AFF_NONE || (scale_a && shift_a), SEMB_ESHAPE, "affine bwd apply sums: missing scale/shift a");
    SEMB_REQUIRE(d->mode_a != SEMB_AFF_BATCH || (mean_a && invstd_a && count_a > 0.f), SEMB_ESHAPE, "affine bwd apply sums: missing batch-norm terms for a");
    SEMB_REQUIRE(!b || d->mode_b == SEMB_AFF_NONE || (scale_b && shift_b), SEMB_ESHAPE, "affine bwd apply sums: missing scale/shift b");
    SEMB_REQUIRE(!b || d->mode_b != SEMB_AFF_BATCH || (mean_b && invstd_b && count_b > 0.f), SEMB_ESHAPE, "affine bwd apply sums: missing batch-norm terms for b");
    AffArgs p{};
    p.N = d->N; p.HW = d->HW; p.C = d->C; p.act = d->act; p.actb = d->actb; p.mode_a = d->mode_a; p.mode_b = d->mode_b;
    p.aff_nstride = d->aff_nstride;
    set_segments(p, d);
    p.a = mkview(a); p.b = mkview(b); p.dy = mkview(dy); p.da = mkview(da); p.db = mkview(db);
    p.scale_a = scale_a; p.shift_a = shift_a; p.mean_a = mean_a; p.invstd_a = invstd_a;
    p.scale_b = scale_b; p.shift_b = shift_b; p.mean_b = mean_b; p.invstd_b = invstd_b;
    p.count_a = count_a; p.count_b = count_b;
    p.dgamma_a = dgamma_a; p.dbeta_a = dbeta_a; p.dgamma_b = dgamma_b; p.dbeta_b = dbeta_b;
    p.stats = const_cast<float*>(sums); p.stats_nstride = sums_nstride; p.stats_cstride = sums_cstride;
    p.acc_a = acc_a; p.acc_b = acc_b;
    int blocks = b ? 2 : 3;
    static const int walk_back = [] { const char* e = getenv("SEMB_AFF_APPLY_FORWARD"); return (e && e[0] && e[0] != '0') ? 0 : 1; }();
    p.rev = walk_back;
    const size_t smem = ring_setup(p, d->dtype, 0, ring_stages(b ? RING_APPB : RING_APP), b ? 3 : 2, &blocks);
    dim3 grid = aff_grid(p, d->aff_nstride != 0 || sums_nstride != 0, blocks);
    cudaStream_t st = as_stream(stream);
    SEMB_AFF_LAUNCH(affine_act_bwd_apply_kernel, d->dtype, b != nullptr, grid, smem, st, p);
    return check_launch("affine_act_bwd_apply_sums");
}

// ---- fused backward ----------------------------------------------------------------------------------------------
template <typename T, bool HAS_B, int ACT, int ACTB>
static cudaError_t launch_fused(dim3 grid, size_t smem, cudaStream_t st, const AffArgs& p, int* max_blocks_per_sm) {
    auto k = affine_act_bwd_fused_kernel<T, HAS_B, ACT, ACTB>;
    if (max_blocks_per_sm) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(max_blocks_per_sm, k, 256, smem);
    }
    void* args[] = {const_cast<AffArgs*>(&p)};
    return cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(k), grid, dim3(256), args, smem, st);
}

template <typename T, bool HAS_B>
static cudaError_t dispatch_fused(dim3 grid, size_t smem, cudaStream_t st, const AffArgs& p, int* occ) {
    if (p.act == SEMB_ACT_NONE && (!HAS_B || p.actb == SEMB_ACT_NONE)) return launch_fused<T, HAS_B, SEMB_ACT_NONE, SEMB_ACT_NONE>(grid, smem, st, p, occ);
    if (p.act == SEMB_ACT_RELU && (!HAS_B || p.actb == SEMB_ACT_NONE)) return launch_fused<T, HAS_B, SEMB_ACT_RELU, SEMB_ACT_NONE>(grid, smem, st, p, occ);
    if (p.act == SEMB_ACT_RELU && p.actb == SEMB_ACT_RELU) return launch_fused<T, HAS_B, SEMB_ACT_RELU, SEMB_ACT_RELU>(grid, smem, st, p, occ);
    return launch_fused<T, HAS_B, -1, -1>(grid, smem, st, p, occ);
}

extern "C" int semb_affine_act_bwd_fused(const semb_affine_desc* d, const semb_tensor* dy, const semb_tensor* a, const semb_tensor* b,
                                         const float* scale_a, const float* shift_a, const float* mean_a, const float* invstd_a,
                                         float count_a, float* dgamma_a, float* dbeta_a,
                                         const float* scale_b, const float* shift_b, const float* mean_b, const float* invstd_b,
                                         float count_b, float* dgamma_b, float* dbeta_b,
                                         float* sums, int32_t sums_nstride, int32_t sums_cstride, void* barrier,
                                         const semb_tensor* da, int32_t acc_a, const semb_tensor* db, int32_t acc_b, void* stream) {
    int rc = check_aff(d);
    if (rc) return rc;
    SEMB_REQUIRE(view_ok(dy) && view_ok(a) && (!b || view_ok(b)), SEMB_EALIGN, "affine bwd fused: bad dy/a/b view");
    SEMB_REQUIRE((!da || view_ok(da)) && (!db || view_ok(db)), SEMB_EALIGN, "affine bwd fused: bad gradient view");
    SEMB_REQUIRE(sums && barrier, SEMB_ESHAPE, "affine bwd fused: null sums / barrier");
    SEMB_REQUIRE(d->mode_a == SEMB_AFF_BATCH || (b && d->mode_b == SEMB_AFF_BATCH), SEMB_ESHAPE,
                 "affine bwd fused: no batch-statistics operand (use semb_affine_act_bwd_apply)");
    SEMB_REQUIRE(d->mode_a == SEMB_AFF_NONE || (scale_a && shift_a), SEMB_ESHAPE, "affine bwd fused: missing scale/shift a");
    SEMB_REQUIRE(d->mode_a != SEMB_AFF_BATCH || (mean_a && invstd_a && count_a > 0.f), SEMB_ESHAPE, "affine bwd fused: missing batch-norm terms for a");
    SEMB_REQUIRE(!b || d->mode_b == SEMB_AFF_NONE || (scale_b && shift_b), SEMB_ESHAPE, "affine bwd fused: missing scale/shift b");
    SEMB_REQUIRE(!b || d->mode_b != SEMB_AFF_BATCH || (mean_b && invstd_b && count_b > 0.f), SEMB_ESHAPE, "affine bwd fused: missing batch-norm terms for b");
    AffArgs p{};
    p.N = d->N; p.HW = d->HW; p.C = d->C; p.act = d->act; p.actb = d->actb; p.mode_a = d->mode_a; p.mode_b = d->mode_b;
    p.aff_nstride = d->aff_nstride;
    set_segments(p, d);
    p.a = mkview(a); p.b = mkview(b); p.dy = mkview(dy); p.da = mkview(da); p.db = mkview(db);
    p.scale_a = scale_a; p.shift_a = shift_a; p.mean_a = mean_a; p.invstd_a = invstd_a;
    p.scale_b = scale_b; p.shift_b = shift_b; p.mean_b = mean_b; p.invstd_b = invstd_b;
    p.count_a = count_a; p.count_b = count_b;
    p.dgamma_a = dgamma_a; p.dbeta_a = dbeta_a; p.dgamma_b = dgamma_b; p.dbeta_b = dbeta_b;
    p.stats = sums; p.stats_nstride = sums_nstride; p.stats_cstride = sums_cstride;
    p.barrier = reinterpret_cast<unsigned int*>(barrier);
    p.acc_a = acc_a; p.acc_b = acc_b;
    const bool has_b = b != nullptr;
    const size_t smem = ((size_t)(256 / (d->C / 8)) * (has_b ? 4 : 2) * d->C + 256) * sizeof(float);
    cudaStream_t st = as_stream(stream);
    // co-residency: the grid must fit one wave (cooperative launch)
    int occ = 0;
    cudaError_t e;
    if (d->dtype == SEMB_BF16) e = has_b ? dispatch_fused<bf16, true>(dim3(), smem, st, p, &occ) : dispatch_fused<bf16, false>(dim3(), smem, st, p, &occ);
    else e = has_b ? dispatch_fused<float, true>(dim3(), smem, st, p, &occ) : dispatch_fused<float, false>(dim3(), smem, st, p, &occ);
    SEMB_REQUIRE(e == cudaSuccess && occ >= 1, SEMB_ECUDA, "affine bwd fused: occupancy query failed (%s)", cudaGetErrorString(e));
    if (occ > (has_b ? 2 : 3)) occ = has_b ? 2 : 3;
    const bool per_sample = d->aff_nstride != 0 || sums_nstride != 0;
    dim3 grid;
    const int rows = 256 / (p.C / 8);
    if (!per_sample) {
        grid = aff_grid(p, false, occ);
    } else {
        long long per_n = (148LL * occ) / p.N;                 // blocks per sample so that the whole grid is one wave
        if (per_n < 1) per_n = 1;
        int ppb = (int)cdivl(p.HW, per_n);
        if (ppb < rows * 4) ppb = rows * 4;
        p.ppb = cdiv(ppb, rows) * rows;
        grid = dim3(cdiv(p.HW, p.ppb), p.N);
    }
    SEMB_REQUIRE((long long)grid.x * grid.y <= 148LL * occ, SEMB_EWORKSPACE,
                 "affine bwd fused: %u x %u blocks do not fit one wave (use the two-pass kernels)", grid.x, grid.y);
    if (d->dtype == SEMB_BF16) e = has_b ? dispatch_fused<bf16, true>(grid, smem, st, p, nullptr) : dispatch_fused<bf16, false>(grid, smem, st, p, nullptr);
    else e = has_b ? dispatch_fused<float, true>(grid, smem, st, p, nullptr) : dispatch_fused<float, false>(grid, smem, st, p, nullptr);
    if (e != cudaSuccess) { set_error("affine bwd fused: cooperative launch failed: %s", cudaGetErrorString(e)); return SEMB_ECUDA; }
    return check_launch("affine_act_bwd_fused");
}

extern "C" int semb_channel_sum(const semb_tensor* x, int32_t N, int32_t HW, float* out, int32_t dtype, void* stream) {
    SEMB_REQUIRE(view_ok(x) && out && N > 0 && HW > 0 && x->C <= 2048, SEMB_ESHAPE, "channel_sum: bad arguments");
    AffArgs p{};
    p.N = N; p.HW = HW; p.C = x->C; p.a = mkview(x); p.stats = out;
    p.ppb = pick_ppb(HW, N, x->C);
    dim3 grid(cdiv(HW, p.ppb), N);
    if (dtype == SEMB_BF16) channel_sum_kernel<bf16><<<grid, 256, x->C * sizeof(float), as_stream(stream)>>>(p);
    else channel_sum_kernel<float><<<grid, 256, x->C * sizeof(float), as_stream(stream)>>>(p);
    return check_launch("channel_sum");
}
