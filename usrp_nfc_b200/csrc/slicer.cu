// slicer.cu -- envelope + ratio classification + gated moving average + hysteresis, sm_100a.
//
// Replaces the per-sample loop of transition_sink.work_stable (transition_sink.py:55-82) and the
// envelope blocks in front of it (decoder.py:25-28, usrp_src.py:31).  Output: the positions where
// the slicer's `val` changes (runs.cu turns those into the reference's event list).
//
// Parallel formulation (validated on the CPU by tests/algomodel.py):
//  * One CTA walks one time segment tile by tile (T = NT*K samples).  The window sum obeys
//    ss[i+1] = ss[i] + admit[i] * (x[i] - ring[i mod L]) where admit depends on ss[i] itself.
//    Because the recurrence is causal, ANY self-consistent assignment of classes equals the
//    sequential answer: guess the classes with ss frozen at the tile start, prefix-sum the gated
//    deltas, re-classify every sample with its own exact ss, repeat until nothing changes.
//  * Sums are carried in double.  While every admitted sample lies within `span_limit` binades
//    all partial sums are exactly representable, so the result does not depend on summation order
//    and equals the reference's running `ss += cur - prev`.  The exponent range is tracked; a
//    segment that leaves it is flagged and redone by the strictly sequential kernel below.
//  * The reference's ratio test  RN64(x*L/ss) vs lo/hi  is decided by the sign of one FMA
//    (exact), with the real division only inside a 2^-50 relative band around the threshold.
//  * "ratio > hi is ignored while cur_state == 2" depends on cur_state only through the
//    distance to the last LOW sample (see vals_from_classes in tests/algomodel.py).
//  * Segments other than the first start cold (ring <- previous L samples, like the
//    reference's warm-up) `halo` samples early; the state they reach at their first emitted
//    sample is compared bit for bit with the predecessor's final state (seam_compare_kernel)
//    and the segment is redone from the true state when it differs.
#include "common.cuh"

namespace nfc {

static const unsigned FULL = 0xffffffffu;
enum { CLS_LOW = -1, CLS_MID = 0, CLS_HIGH = 1 };

// ---------------------------------------------------------------- sample loading / envelope
__device__ __forceinline__ float env_real(float s) { return __fmul_rn(s, s); }  // float_to_complex + mag^2, im = 0
__device__ __forceinline__ float env_iq(float re, float im) { return __fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im)); }

__device__ __forceinline__ float load_one(const void *in, int64_t idx, const SlicerParams &p) {
    switch (p.input_kind) {
        case IN_ENVELOPE_F32: return __ldg(reinterpret_cast<const float *>(in) + idx);
        case IN_REAL_F32: return env_real(__ldg(reinterpret_cast<const float *>(in) + idx));
        case IN_IQ_F32: {
            float2 c = __ldg(reinterpret_cast<const float2 *>(in) + idx);
            return env_iq(c.x, c.y);
        }
        default: {
            float s = __fdiv_rn((float)__ldg(reinterpret_cast<const short *>(in) + idx), p.pcm_scale);
            return env_real(s);
        }
    }
}

// streaming 128-bit load: the sample stream is read exactly once
__device__ __forceinline__ float4 ldg_stream4(const float4 *ptr) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(ptr));
    return r;
}

template <int K>
__device__ __forceinline__ void load_samples(const SegWork &w, const SlicerParams &p, int64_t p0, float (&x)[K]) {
    const int64_t i0 = p0 - w.in_pos0;
    if (K == 4 && p0 >= w.in_begin && p0 + 4 <= w.in_end) {
        if (p.input_kind == IN_ENVELOPE_F32 || p.input_kind == IN_REAL_F32) {
            float4 v = ldg_stream4(reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(w.in) + i0));
            x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
            if (p.input_kind == IN_REAL_F32) {
#pragma unroll
                for (int j = 0; j < K; j++) x[j] = env_real(x[j]);
            }
            return;
        }
        if (p.input_kind == IN_IQ_F32) {
            const float4 *q = reinterpret_cast<const float4 *>(reinterpret_cast<const float2 *>(w.in) + i0);
            float4 a = ldg_stream4(q), b = ldg_stream4(q + 1);
            x[0] = env_iq(a.x, a.y); x[1] = env_iq(a.z, a.w); x[2] = env_iq(b.x, b.y); x[3] = env_iq(b.z, b.w);
            return;
        }
        if (p.input_kind == IN_PCM_S16) {
            short4 s = __ldg(reinterpret_cast<const short4 *>(reinterpret_cast<const short *>(w.in) + i0));
            x[0] = env_real(__fdiv_rn((float)s.x, p.pcm_scale));
            x[1] = env_real(__fdiv_rn((float)s.y, p.pcm_scale));
            x[2] = env_real(__fdiv_rn((float)s.z, p.pcm_scale));
            x[3] = env_real(__fdiv_rn((float)s.w, p.pcm_scale));
            return;
        }
    }
#pragma unroll
    for (int j = 0; j < K; j++) {
        int64_t q = p0 + j;
        x[j] = (q >= w.in_begin && q < w.in_end) ? load_one(w.in, q - w.in_pos0, p) : 0.0f;
    }
}

// ---------------------------------------------------------------- the ratio test, exactly
// transition_sink.py:59-71:  ratio = bit*length/ss  (or the ss == 0 constants);  lo > ratio -> LOW;
// ratio > hi -> HIGH (before hysteresis); else MID.
__device__ __noinline__ int classify_div(double pr, double ss, double lo, double hi) {
    double ratio = __ddiv_rn(pr, ss);
    if (lo > ratio) return CLS_LOW;
    if (ratio > hi) return CLS_HIGH;
    return CLS_MID;
}

__device__ __forceinline__ int classify(float x, double ss, const SlicerParams &p) {
    if (ss == 0.0) return x == 0.0f ? p.cls_ss0_x0 : p.cls_ss0_xn;
    const double pr = (double)x * p.Ld;  // exact: 24-bit significand times an integer < 2^29
    if (!(ss > 0.0) || !(p.lo > 0.0) || !(p.hi > 0.0)) return classify_div(pr, ss, p.lo, p.hi);
    // sign(pr - lo*ss) is exact with one FMA; RN(pr/ss) can only disagree with it within half an ulp
    const double r = fma(-p.lo, ss, pr);
    if (r < 0.0) {
        if (r > -(p.lo * ss) * 0x1p-50) return classify_div(pr, ss, p.lo, p.hi);
        return CLS_LOW;
    }
    const double r2 = fma(-p.hi, ss, pr);
    if (r2 > 0.0) {
        if (r2 < (p.hi * ss) * 0x1p-50) return classify_div(pr, ss, p.lo, p.hi);
        return CLS_HIGH;
    }
    // r >= 0 and r2 <= 0: pr/ss in [lo, hi] exactly, and RN is monotonic
    return CLS_MID;
}

// hysteresis: is a HIGH-class sample at stream index i forced to val 0 (cur_state == 2)?
// b = last LOW sample before i, a = first sample of the LOW run containing b.
__device__ __forceinline__ bool st2_forced(int64_t i, int64_t b, int64_t a, int mx) {
    if (b == NO_POS || i - b > (int64_t)mx + 1) return false;
    const int64_t j = b - a;
    const bool tmo = j >= mx && (j % mx) == 0;  // b itself was a timeout sample: cur_state went back to 0
    return !tmo;
}

// ---------------------------------------------------------------- block-wide primitives
template <int NT>
struct BlockShared {
    static const int NW = NT / 32;
    double wsum[2][NW];
    int wcnt[2][NW];
    int wlastcls[NW];   // class of each warp's last sample (current round)
    int wmaxL[NW];      // tile-relative index of the last LOW sample per warp
    int wmaxS[NW];      // tile-relative index of the last LOW-run start per warp
    unsigned flags[3];
    int last_val;
    int emin, emax;
    double red[NW];
};

template <int NT>
__device__ __forceinline__ double block_excl_scan(double v, double &total, double (*wsum)[NT / 32], int buf) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double n = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) wsum[buf][warp] = inc;
    __syncthreads();
    double wbase = 0.0, tot = 0.0;
#pragma unroll
    for (int w = 0; w < NT / 32; w++) {
        double s = wsum[buf][w];
        if (w < warp) wbase += s;
        tot += s;
    }
    total = tot;
    return wbase + (inc - v);
}

template <int NT>
__device__ __forceinline__ int block_excl_scan_int(int v, int &total, int (*wcnt)[NT / 32], int buf) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int n = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) wcnt[buf][warp] = inc;
    __syncthreads();
    int wbase = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; w++) {
        int s = wcnt[buf][w];
        if (w < warp) wbase += s;
        tot += s;
    }
    total = tot;
    return wbase + (inc - v);
}

__device__ __forceinline__ void exp_track(float x, int &emin, int &emax) {
    const unsigned ub = __float_as_uint(x);
    if (ub != 0u) {
        const int e = (int)(ub >> 23);  // sign bit set -> e >= 256 -> flagged not sane below
        emin = min(emin, max(e, 1));
        emax = max(emax, e);
    }
}

// ---------------------------------------------------------------- the segment kernel
template <int NT, int K>
__global__ void __launch_bounds__(NT) slicer_kernel(const SegWork *__restrict__ works,
                                                    const SlicerParams *__restrict__ params) {
    constexpr int T = NT * K;
    constexpr int NW = NT / 32;
    extern __shared__ __align__(16) float ring[];
    __shared__ BlockShared<NT> sh;

    const SegWork w = works[blockIdx.x];
    const SlicerParams p = params[w.param_idx];
    const int L = p.L, mx = p.mx;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- entry state
    double ss0;
    int64_t lastL, lrun_start;
    int last_val;
    int status = SEG_OK;
    int emin = 1 << 30, emax = 0;

    if (w.state_in) {
        const float *src = state_ring(w.state_in);
        for (int i = tid; i < L; i += NT) {
            float v = src[i];
            ring[i] = v;
            exp_track(v, emin, emax);
        }
        ss0 = w.state_in->ss;
        lastL = w.state_in->lastL;
        lrun_start = w.state_in->lrun_start;
        last_val = w.state_in->last_val;
    } else {
        // cold start: the previous L samples, unconditionally (transition_sink.py:118)
        double part = 0.0;
        for (int i = tid; i < L; i += NT) {
            const int64_t q = w.warm_begin - L + i;
            float v = load_one(w.in, q - w.in_pos0, p);
            ring[(int)(q % L)] = v;
            part += (double)v;
            exp_track(v, emin, emax);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(FULL, part, o);
        if (lane == 0) sh.red[warp] = part;
        __syncthreads();
        ss0 = 0.0;
        for (int i = 0; i < NW; i++) ss0 += sh.red[i];
        lastL = NO_POS;
        lrun_start = NO_POS;
        last_val = 0;
    }
    if (tid == 0) {
        sh.flags[0] = sh.flags[1] = sh.flags[2] = 0u;
        sh.last_val = last_val;
    }
    __syncthreads();

    uint32_t seg_count = 0;
    int scan_buf = 0, cnt_buf = 0;
    unsigned round_no = 0;

    const int64_t tile_first = w.warm_begin / T;
    const int64_t tile_last = (w.end > w.warm_begin) ? (w.end - 1) / T : tile_first - 1;
    int slot0 = (int)((tile_first * T + (int64_t)tid * K) % L);

    for (int64_t tile = tile_first; tile <= tile_last; tile++) {
        const int64_t P0 = tile * T;
        const int64_t p0 = P0 + (int64_t)tid * K;

        if (w.seam_in && P0 == w.begin && w.begin > w.warm_begin) {
            // snapshot of the speculative state at the first emitted sample
            float *dst = state_ring(w.seam_in);
            for (int i = tid; i < L; i += NT) dst[i] = ring[i];
            if (tid == 0) {
                SlicerHdr h;
                h.ss = ss0; h.pos = w.begin; h.lastL = lastL; h.lrun_start = lrun_start;
                h.last_val = last_val; h.emin = 0; h.emax = 0; h.status = 0;
                *w.seam_in = h;
            }
        }

        float x[K], prev[K];
        bool act[K];
        load_samples<K>(w, p, p0, x);
#pragma unroll
        for (int j = 0; j < K; j++) act[j] = (p0 + j >= w.warm_begin) && (p0 + j < w.end);

        // ring slots of this thread's samples: (p0 + j) mod L
        int slot[K];
#pragma unroll
        for (int j = 0; j < K; j++) {
            int s = slot0 + j;
            slot[j] = s >= L ? s - L : s;
        }
        if (K == 4 && (L & 3) == 0) {
            float4 v = *reinterpret_cast<const float4 *>(ring + slot0);
            prev[0] = v.x; prev[1] = v.y; prev[2] = v.z; prev[3] = v.w;
        } else {
#pragma unroll
            for (int j = 0; j < K; j++) prev[j] = ring[slot[j]];
        }

        double dl[K];
        int cls[K];
        bool forced[K];
#pragma unroll
        for (int j = 0; j < K; j++) {
            dl[j] = act[j] ? (double)x[j] - (double)prev[j] : 0.0;  // cur - prev (transition_sink.py:82)
            cls[j] = act[j] ? classify(x[j], ss0, p) : CLS_MID;      // guess: ss frozen at the tile start
            forced[j] = false;
        }
        const bool carry_recent = lastL != NO_POS && (P0 - lastL) <= (int64_t)mx + 1;

        double total = 0.0;
        bool had_slow = false, anyL = false;
        // ---- fix-point: classes <-> prefix sums of the gated deltas
        for (;;) {
            const unsigned fb = round_no % 3u;
            if (tid == 0) sh.flags[(round_no + 1u) % 3u] = 0u;
            double pre[K];
            double run = 0.0;
#pragma unroll
            for (int j = 0; j < K; j++) {
                pre[j] = run;
                const bool admit = act[j] && (cls[j] == CLS_MID || (cls[j] == CLS_HIGH && forced[j]));
                if (admit) run += dl[j];
            }
            const double base = block_excl_scan<NT>(run, total, sh.wsum, scan_buf);
            scan_buf ^= 1;

            unsigned f = 0u;
#pragma unroll
            for (int j = 0; j < K; j++) {
                if (act[j]) {
                    const int c = classify(x[j], ss0 + (base + pre[j]), p);
                    if (c != cls[j]) { f |= 1u; cls[j] = c; }
                    if (c == CLS_HIGH) f |= 2u;
                    if (c == CLS_LOW) f |= 4u;
                }
            }
            f = __reduce_or_sync(FULL, f);
            if (lane == 0 && f) atomicOr(&sh.flags[fb], f);
            if (lane == 31) sh.wlastcls[warp] = act[K - 1] ? cls[K - 1] : 2;  // 2 = "no sample"
            __syncthreads();
            unsigned flags = sh.flags[fb];
            round_no++;

            if ((flags & 2u) && ((flags & 4u) || carry_recent)) {
                // ---- hysteresis: distance from each HIGH sample to the last LOW sample / its run start
                // previous sample's class for this thread's first sample
                int pc = __shfl_up_sync(FULL, cls[K - 1], 1);
                const bool pact = __shfl_up_sync(FULL, (int)act[K - 1], 1) != 0;
                if (lane == 0) {
                    pc = CLS_MID;
                    bool found = false;
                    for (int ww = warp - 1; ww >= 0 && !found; ww--) {
                        int c = sh.wlastcls[ww];
                        if (c != 2) { pc = c; found = true; }
                    }
                    if (!found) pc = (last_val == -1) ? CLS_LOW : CLS_MID;
                } else if (!pact) {
                    pc = (last_val == -1) ? CLS_LOW : CLS_MID;  // inactive prefix of the first tile
                }
                // thread-local last LOW index / last LOW-run start index (tile relative), inclusive of own samples
                int myL = -1, myS = -1;
                int preL[K], preS[K];
                {
                    int c_prev = pc;
#pragma unroll
                    for (int j = 0; j < K; j++) {
                        preL[j] = myL;
                        preS[j] = myS;
                        if (act[j]) {
                            if (cls[j] == CLS_LOW) {
                                if (c_prev != CLS_LOW) myS = tid * K + j;
                                myL = tid * K + j;
                            }
                            c_prev = cls[j];
                        }
                    }
                }
                int incL = myL, incS = myS;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    int a = __shfl_up_sync(FULL, incL, o), b = __shfl_up_sync(FULL, incS, o);
                    if (lane >= o) { incL = max(incL, a); incS = max(incS, b); }
                }
                int exL = __shfl_up_sync(FULL, incL, 1), exS = __shfl_up_sync(FULL, incS, 1);
                if (lane == 0) { exL = -1; exS = -1; }
                if (lane == 31) { sh.wmaxL[warp] = incL; sh.wmaxS[warp] = incS; }
                __syncthreads();
                for (int ww = 0; ww < warp; ww++) {
                    exL = max(exL, sh.wmaxL[ww]);
                    exS = max(exS, sh.wmaxS[ww]);
                }
                unsigned f2 = 0u;
#pragma unroll
                for (int j = 0; j < K; j++) {
                    bool fo = false;
                    if (act[j] && cls[j] == CLS_HIGH) {
                        const int bl = max(exL, preL[j]);
                        const int sl = max(exS, preS[j]);
                        const int64_t b = bl >= 0 ? P0 + bl : lastL;
                        const int64_t a = sl >= 0 ? P0 + sl : lrun_start;
                        fo = st2_forced(p0 + j, b, a, mx);
                    }
                    if (fo != forced[j]) { f2 = 1u; forced[j] = fo; }
                }
                // a second vote: did any `forced` flag change?
                const unsigned fb2 = round_no % 3u;
                if (tid == 0) sh.flags[(round_no + 1u) % 3u] = 0u;
                f2 = __reduce_or_sync(FULL, f2);
                if (lane == 0 && f2) atomicOr(&sh.flags[fb2], f2);
                __syncthreads();
                flags |= sh.flags[fb2] & 1u;
                round_no++;
                had_slow = true;
            } else if (had_slow) {
                // `forced` flags are only ever set in the (block-uniform) branch above; once the tile no
                // longer needs it they are cleared, which changes `admit`, so vote for another round.
                unsigned f2 = 0u;
#pragma unroll
                for (int j = 0; j < K; j++) {
                    if (forced[j]) { f2 = 1u; forced[j] = false; }
                }
                if (__syncthreads_or((int)f2)) flags |= 1u;
                had_slow = false;
            }
            if (!(flags & 1u)) {  // converged: `total` belongs to the final classes
                anyL = (flags & 4u) != 0u;
                break;
            }
        }

        // ---- ring update (transition_sink.py:75-81) and exponent tracking
        float nv[K];
#pragma unroll
        for (int j = 0; j < K; j++) {
            const bool admit = act[j] && (cls[j] == CLS_MID || (cls[j] == CLS_HIGH && forced[j]));
            nv[j] = admit ? x[j] : prev[j];
            if (admit) exp_track(x[j], emin, emax);
        }
        if (K == 4 && (L & 3) == 0 && act[0] && act[K - 1]) {
            *reinterpret_cast<float4 *>(ring + slot0) = make_float4(nv[0], nv[1], nv[2], nv[3]);
        } else {
#pragma unroll
            for (int j = 0; j < K; j++)
                if (act[j]) ring[slot[j]] = nv[j];
        }

        // ---- vals and transitions
        int val[K];
#pragma unroll
        for (int j = 0; j < K; j++)
            val[j] = act[j] ? (cls[j] == CLS_LOW ? -1 : ((cls[j] == CLS_HIGH && !forced[j]) ? 1 : 0)) : 3;  // 3 = none
        // last defined val of this thread, and of the threads before it
        int mylast = 3;
#pragma unroll
        for (int j = 0; j < K; j++)
            if (val[j] != 3) mylast = val[j];
        // previous val for the first sample: nearest earlier thread with a defined val, else the carry
        int pv;
        {
            // warp-level: find nearest lower lane with mylast != 3
            const unsigned has = __ballot_sync(FULL, mylast != 3);
            const unsigned lower = has & ((1u << lane) - 1u);
            const int src = lower ? 31 - __clz(lower) : 0;
            const int got = __shfl_sync(FULL, mylast, src);
            pv = lower ? got : 3;
            int wl = __shfl_sync(FULL, mylast, has ? 31 - __clz(has) : 0);
            if (!has) wl = 3;
            if (lane == 0) sh.wlastcls[warp] = wl;  // reuse: last defined val of the warp (3 = none)
        }
        __syncthreads();
        if (pv == 3) {
            pv = last_val;
            for (int ww = warp - 1; ww >= 0; ww--) {
                const int c = sh.wlastcls[ww];
                if (c != 3) { pv = c; break; }
            }
        }
        int ntr = 0;
        unsigned trmask = 0u;
        int maxS = -1, maxL = -1;
        {
            int c_prev = pv;
#pragma unroll
            for (int j = 0; j < K; j++) {
                if (val[j] != 3) {
                    if (val[j] != c_prev) {
                        if (p0 + j >= w.begin) { trmask |= 1u << j; ntr++; }
                        if (val[j] == -1) maxS = tid * K + j;
                    }
                    if (val[j] == -1) maxL = tid * K + j;
                    c_prev = val[j];
                }
            }
        }
        if (anyL) {
            // carry of (lastL, lrun_start): block maxima
            int mL = __reduce_max_sync(FULL, maxL), mS = __reduce_max_sync(FULL, maxS);
            if (lane == 0) { sh.wmaxL[warp] = mL; sh.wmaxS[warp] = mS; }
        }
        int tot_tr = 0;
        const int tr_base = block_excl_scan_int<NT>(ntr, tot_tr, sh.wcnt, cnt_buf);
        cnt_buf ^= 1;
        if (tot_tr) {
            uint32_t idx = seg_count + (uint32_t)tr_base;
#pragma unroll
            for (int j = 0; j < K; j++) {
                if (trmask & (1u << j)) {
                    if (idx < w.trans_cap) w.trans[idx] = pack_trans((uint32_t)(p0 + j - w.slab_pos0), val[j]);
                    idx++;
                }
            }
            seg_count += (uint32_t)tot_tr;
        }
        // ---- carries into the next tile
        ss0 += total;
        if (anyL) {
            int mL = -1, mS = -1;
#pragma unroll
            for (int ww = 0; ww < NW; ww++) {
                mL = max(mL, sh.wmaxL[ww]);
                mS = max(mS, sh.wmaxS[ww]);
            }
            if (mL >= 0) {
                lastL = P0 + mL;
                if (mS >= 0) lrun_start = P0 + mS;
            }
        }
        {
            // last defined val of the tile
            int lv = 3;
            for (int ww = NW - 1; ww >= 0; ww--) {
                const int c = sh.wlastcls[ww];
                if (c != 3) { lv = c; break; }
            }
            if (lv != 3) last_val = lv;
        }
        slot0 += T;
        if (slot0 >= L) slot0 -= L;
        __syncthreads();  // ring writes and shared scratch settle before the next tile reads them
    }

    // ---- exit: exactness audit and final state
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        emin = min(emin, __shfl_xor_sync(FULL, emin, o));
        emax = max(emax, __shfl_xor_sync(FULL, emax, o));
    }
    if (tid == 0) { sh.emin = 1 << 30; sh.emax = 0; }
    __syncthreads();
    if (lane == 0) { atomicMin(&sh.emin, emin); atomicMax(&sh.emax, emax); }
    __syncthreads();
    emin = sh.emin; emax = sh.emax;
    if (emax >= 255) status |= SEG_NOT_SANE;
    if (emax > 0 && emax - emin > p.span_limit) status |= SEG_INEXACT;
    if (seg_count > w.trans_cap) status |= SEG_OVERFLOW;

    if (w.state_out) {
        float *dst = state_ring(w.state_out);
        for (int i = tid; i < L; i += NT) dst[i] = ring[i];
        if (tid == 0) {
            SlicerHdr h;
            h.ss = ss0; h.pos = w.end; h.lastL = lastL; h.lrun_start = lrun_start;
            h.last_val = last_val; h.emin = emin; h.emax = emax; h.status = status;
            *w.state_out = h;
        }
    }
    if (tid == 0 && w.trans_count) *w.trans_count = seg_count;
    if (tid == 0 && w.status) *w.status = status;
}

// ---------------------------------------------------------------- strictly sequential path
// The reference recurrence, literally, one thread per segment: used when av_window is smaller
// than a tile, when the exponent audit fails (inexact sums, negative or non-finite samples) and
// as the independent on-device cross-check of the parallel kernel.  `ss` is carried exactly as
// the reference carries it (ss += cur - prev in double, in order).
__global__ void slicer_serial_kernel(const SegWork *__restrict__ works, const SlicerParams *__restrict__ params,
                                     float *__restrict__ ring_scratch, size_t ring_stride) {
    if (threadIdx.x != 0) return;
    const SegWork w = works[blockIdx.x];
    const SlicerParams p = params[w.param_idx];
    const int L = p.L, mx = p.mx;
    float *ring = ring_scratch + (size_t)blockIdx.x * ring_stride;

    double ss;
    int64_t lastL, lrun_start;
    int last_val;
    if (w.state_in) {
        const float *src = state_ring(w.state_in);
        for (int i = 0; i < L; i++) ring[i] = src[i];
        ss = w.state_in->ss;
        lastL = w.state_in->lastL;
        lrun_start = w.state_in->lrun_start;
        last_val = w.state_in->last_val;
    } else {
        ss = 0.0;
        // slot order == stream order here only if warm_begin % L == 0; sum in stream order like sum(ar)
        for (int i = 0; i < L; i++) {
            const int64_t q = w.warm_begin - L + i;
            float v = load_one(w.in, q - w.in_pos0, p);
            ring[(int)(q % L)] = v;
            ss += (double)v;
        }
        lastL = NO_POS;
        lrun_start = NO_POS;
        last_val = 0;
    }
    uint32_t count = 0;
    int slot = (int)(w.warm_begin % L);
    for (int64_t q = w.warm_begin; q < w.end; q++) {
        if (w.seam_in && q == w.begin && w.begin > w.warm_begin) {
            float *dst = state_ring(w.seam_in);
            for (int i = 0; i < L; i++) dst[i] = ring[i];
            SlicerHdr h;
            h.ss = ss; h.pos = q; h.lastL = lastL; h.lrun_start = lrun_start;
            h.last_val = last_val; h.emin = 0; h.emax = 0; h.status = 0;
            *w.seam_in = h;
        }
        const float x = load_one(w.in, q - w.in_pos0, p);
        const float prev = ring[slot];
        int c;
        if (ss == 0.0) c = x == 0.0f ? p.cls_ss0_x0 : p.cls_ss0_xn;
        else c = classify_div((double)x * p.Ld, ss, p.lo, p.hi);
        int val;
        if (c == CLS_LOW) {
            val = -1;
            if (last_val != -1) lrun_start = q;
            lastL = q;
        } else if (c == CLS_HIGH && !st2_forced(q, lastL, lrun_start, mx)) {
            val = 1;
        } else {
            val = 0;
            ring[slot] = x;
            ss += ((double)x - (double)prev);
        }
        if (val != last_val) {
            if (q >= w.begin) {
                if (count < w.trans_cap) w.trans[count] = pack_trans((uint32_t)(q - w.slab_pos0), val);
                count++;
            }
            last_val = val;
        }
        slot++;
        if (slot == L) slot = 0;
    }
    if (w.state_out) {
        float *dst = state_ring(w.state_out);
        for (int i = 0; i < L; i++) dst[i] = ring[i];
        SlicerHdr h;
        h.ss = ss; h.pos = w.end; h.lastL = lastL; h.lrun_start = lrun_start;
        h.last_val = last_val; h.emin = 0; h.emax = 0;
        h.status = count > w.trans_cap ? SEG_OVERFLOW : SEG_OK;
        *w.state_out = h;
    }
    if (w.trans_count) *w.trans_count = count;
    if (w.status) *w.status = count > w.trans_cap ? SEG_OVERFLOW : SEG_OK;
}

// ---------------------------------------------------------------- seam verification
// For seam k (between segment k-1 and k): does the state the predecessor really reached equal the
// state the successor assumed at its first emitted sample?  Bitwise on ring and ss; the hysteresis
// carry is compared in canonical form (only what can still influence the future).
__device__ __forceinline__ void canon_st2(const SlicerHdr &h, int mx, int &kind, int64_t &val) {
    if (h.last_val == -1) { kind = 1; val = h.lrun_start; return; }
    if (h.lastL == NO_POS) { kind = 0; val = 0; return; }
    const int64_t j = h.lastL - h.lrun_start;
    const bool tmo = j >= mx && (j % mx) == 0;
    const int64_t until = tmo ? h.lastL : h.lastL + mx + 1;
    if (until < h.pos) { kind = 0; val = 0; } else { kind = 2; val = until; }
}

__global__ void seam_compare_kernel(const SlicerHdr *const *__restrict__ truth, const SlicerHdr *const *__restrict__ assumed,
                                    const int *__restrict__ param_idx, const SlicerParams *__restrict__ params,
                                    int *__restrict__ mismatch) {
    const SlicerHdr *a = truth[blockIdx.x], *b = assumed[blockIdx.x];
    if (!a || !b) return;
    const SlicerParams p = params[param_idx[blockIdx.x]];
    int bad = 0;
    const uint32_t *ra = reinterpret_cast<const uint32_t *>(state_ring(a));
    const uint32_t *rb = reinterpret_cast<const uint32_t *>(state_ring(b));
    for (int i = threadIdx.x; i < p.L; i += blockDim.x) bad |= (ra[i] != rb[i]);
    if (threadIdx.x == 0) {
        if (__double_as_longlong(a->ss) != __double_as_longlong(b->ss)) bad = 1;
        if (a->last_val != b->last_val || a->pos != b->pos) bad = 1;
        int ka, kb;
        int64_t va, vb;
        canon_st2(*a, p.mx, ka, va);
        canon_st2(*b, p.mx, kb, vb);
        if (ka != kb || va != vb) bad = 1;
    }
    if (__syncthreads_or(bad) && threadIdx.x == 0) mismatch[blockIdx.x] = 1;
}

// ---------------------------------------------------------------- host launchers
int launch_slicer(const SegWork *d_works, int n_works, const SlicerParams *d_params, int L, bool vec_ok,
                  cudaStream_t stream) {
    if (n_works <= 0) return 0;
    const size_t smem = ((size_t)L * 4 + 15) / 16 * 16;
    if (vec_ok && L >= 1024) {
        auto k = slicer_kernel<256, 4>;
        NFC_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<n_works, 256, smem, stream>>>(d_works, d_params);
    } else {
        auto k = slicer_kernel<256, 1>;
        NFC_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<n_works, 256, smem, stream>>>(d_works, d_params);
    }
    NFC_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int launch_slicer_serial(const SegWork *d_works, int n_works, const SlicerParams *d_params, float *d_ring_scratch,
                         size_t ring_stride, cudaStream_t stream) {
    if (n_works <= 0) return 0;
    slicer_serial_kernel<<<n_works, 32, 0, stream>>>(d_works, d_params, d_ring_scratch, ring_stride);
    NFC_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int launch_seam_compare(const SlicerHdr *const *d_truth, const SlicerHdr *const *d_assumed, const int *d_param_idx,
                        const SlicerParams *d_params, int *d_mismatch, int n, cudaStream_t stream) {
    if (n <= 0) return 0;
    seam_compare_kernel<<<n, 256, 0, stream>>>(d_truth, d_assumed, d_param_idx, d_params, d_mismatch);
    NFC_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace nfc
