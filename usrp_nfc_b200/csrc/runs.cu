// runs.cu -- val transitions -> the reference's event list, one thread per run.
//
// Replaces the run-length / timeout / emit part of transition_sink.work_stable
// (transition_sink.py:84-99).  The reference carries (cur_state, last_bit, dur) from sample to
// sample; here every field of every event is a closed-form function of the run it closes, the
// run before it and the carry at the window start (derivation and CPU proof: tests/algomodel.py,
// events_from_transitions).  Events are written in stream order (count -> scan -> write).
#include "common.cuh"
#include "scan.cuh"

namespace nfc {

// ---- segment lists -> one dense, ordered transition array ---------------------------------
__global__ void seg_offsets_kernel(const uint32_t *__restrict__ counts, uint32_t *__restrict__ offsets, int n) {
    // n is small (number of segments in a slab): one block, chunked
    __shared__ uint32_t sh[1024];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        uint32_t v = i < n ? counts[i] : 0u;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            uint32_t a = threadIdx.x >= o ? sh[threadIdx.x - o] : 0u;
            __syncthreads();
            sh[threadIdx.x] += a;
            __syncthreads();
        }
        if (i < n) offsets[i] = carry + sh[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += sh[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[n] = carry;
}

__global__ void gather_trans_kernel(const SegWork *__restrict__ works, const uint32_t *__restrict__ counts,
                                    const uint32_t *__restrict__ offsets, TransRec *__restrict__ dense) {
    const SegWork w = works[blockIdx.x];
    const uint32_t n = min(counts[blockIdx.x], w.trans_cap), off = offsets[blockIdx.x];
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) dense[off + i] = w.trans[i];
}

// pieces of segment transition lists -> dense array (offsets computed on the host: there are few pieces)
__global__ void gather_pieces_kernel(const TransRec *const *__restrict__ src, const uint32_t *__restrict__ n,
                                     const uint32_t *__restrict__ offsets, TransRec *__restrict__ dense) {
    const TransRec *s = src[blockIdx.x];
    const uint32_t cnt = n[blockIdx.x], off = offsets[blockIdx.x];
    for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) dense[off + i] = s[i];
}

// ---- per-run event derivation ---------------------------------------------------------------
struct RunView {
    const TransRec *tr;  // dense transitions of the window, ascending
    uint32_t R;          // number of transitions (runs = R + 1; run 0 is the carried-in run)
    int32_t w0, w1;      // window, relative to the slab origin (a slab is at most 2^30 samples: 32-bit positions)
    int st0, lb0, dur0;  // reference's (cur_state, last_bit, dur) at w0
    int mx;

    __device__ __forceinline__ int u(uint32_t r) const { return r == 0 ? lb0 : trans_val(tr[r - 1]); }
    __device__ __forceinline__ int32_t p(uint32_t r) const { return r == 0 ? w0 - dur0 : (int32_t)trans_pos(tr[r - 1]); }
    __device__ __forceinline__ int32_t e(uint32_t r) const { return r < R ? (int32_t)trans_pos(tr[r]) : w1; }

    // run lengths stay below 2^31 (a slab is at most 2^30 samples, the carried-in run adds at most max_len): all the
    // divisions by max_len are 32-bit
    __device__ __forceinline__ bool tmo_at_last(int32_t pp, int32_t ee) const {
        const uint32_t ell1 = (uint32_t)(ee - pp - 1);
        return ell1 >= (uint32_t)mx && (ell1 % (uint32_t)mx) == 0u;
    }
    // first timeout position q = p + k*mx (k >= 1), q >= w0, or -1 when none falls before e
    __device__ __forceinline__ int32_t first_tmo(uint32_t r, int32_t pp, int32_t ee) const {
        int32_t k = 1;
        if (r == 0) {
            const uint32_t need = (uint32_t)(w0 - pp);  // = dur0
            k = (int32_t)((need + (uint32_t)mx - 1u) / (uint32_t)mx);
            if (k < 1) k = 1;
        }
        const int32_t q = pp + k * mx;
        return q < ee ? q : -1;
    }
    // cur_state after the last sample of a run with val != 0
    __device__ __forceinline__ int s_end_nz(uint32_t r, int uu, int32_t pp, int32_t ee) const {
        if (r == 0 && ee == w0) return st0;
        return tmo_at_last(pp, ee) ? 0 : (uu == -1 ? 2 : 1);
    }
    __device__ int s_end(uint32_t r) const {
        const int uu = u(r);
        const int32_t pp = p(r), ee = e(r);
        if (r == 0 && ee == w0) return st0;
        if (uu != 0) return s_end_nz(r, uu, pp, ee);
        if (first_tmo(r, pp, ee) >= 0) return 0;
        if (r == 0) return st0;
        return s_end_nz(r - 1, u(r - 1), p(r - 1), e(r - 1));  // adjacent runs differ, so run r-1 has val != 0
    }
};

// Enumerate the events of run r in order; Emit(rel_pos, v, d, type).
template <class Emit>
__device__ __forceinline__ void run_events(const RunView &V, uint32_t r, bool keep_dropped, Emit emit) {
    const int uu = V.u(r);
    const int32_t pp = V.p(r), ee = V.e(r);
    const int s_begin = r == 0 ? V.st0 : V.s_end(r - 1);
    int32_t q = V.first_tmo(r, pp, ee);
    bool first = true;
    bool any_tmo = false;
    while (q >= 0 && q < ee) {  // timeouts (transition_sink.py:95-99)
        const int st_q = uu == -1 ? 2 : (uu == 1 ? 1 : (first ? s_begin : 0));
        first = false;
        any_tmo = true;
        if (keep_dropped || st_q != 0) emit((uint32_t)q, uu + (st_q == 2 ? 1 : 0), V.mx, st_q - 1);
        if (!keep_dropped && uu == 0) break;  // later timeouts of a 0-run all have cur_state 0 (dropped)
        q += V.mx;
    }
    if (r < V.R) {  // the transition that closes the run (transition_sink.py:86-92)
        int se;
        if (r == 0 && ee == V.w0) se = V.st0;
        else if (uu != 0) se = V.s_end_nz(r, uu, pp, ee);
        else se = any_tmo ? 0 : s_begin;
        const int un = trans_val(V.tr[r]);
        const int st_after = un == -1 ? 2 : (un == 1 ? 1 : se);
        const uint32_t ell1 = (uint32_t)(ee - pp - 1);
        const int d = se == 0 ? V.mx : (int)(ell1 % (uint32_t)V.mx) + 1;
        if (keep_dropped || st_after != 0) emit((uint32_t)ee, uu + (st_after == 2 ? 1 : 0), d, st_after - 1);
    }
}

// What the kernels know about a slab before it runs: the number of transitions and the carry are read from device memory
// (left there by the extraction and by the slab before), so that the host can queue a slab's whole chain of kernels without
// waiting for any count.  cap_R sizes the grid and the buffers: more transitions than that raise POST_OVF_TRANS and the
// host does the slab again with exact sizes.
struct RunArgs {
    const TransRec *tr;
    const uint32_t *d_R;     // transitions of the slab
    uint32_t cap_R;
    int32_t w0, w1;
    const RunCarry *rc_in;   // the reference's (cur_state, last_bit, dur) at w0
    int mx;
    uint32_t *flags;
};
__device__ __forceinline__ RunView make_view(const RunArgs &A) {
    RunView V;
    const RunCarry c = *A.rc_in;
    V.tr = A.tr;
    V.R = min(*A.d_R, A.cap_R);
    V.w0 = A.w0; V.w1 = A.w1;
    V.st0 = c.st; V.lb0 = c.last_bit; V.dur0 = c.dur;
    V.mx = A.mx;
    return V;
}

// counts[r] for every r <= cap_R (0 behind the slab's last run: the scan runs over the capacity)
__global__ void run_count_kernel(RunArgs A, int keep_dropped, uint32_t *__restrict__ counts) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > A.cap_R) return;
    const RunView V = make_view(A);
    if (r == 0 && *A.d_R > A.cap_R) atomicOr(A.flags, (uint32_t)POST_OVF_TRANS);
    uint32_t c = 0;
    if (r <= V.R) run_events(V, r, keep_dropped != 0, [&](uint32_t, int, int, int) { c++; });
    counts[r] = c;
}

__global__ void run_write_kernel(RunArgs A, int keep_dropped, const uint32_t *__restrict__ offsets,
                                 EventRec *__restrict__ out, uint32_t cap, const uint32_t *__restrict__ d_M,
                                 RunCarry *__restrict__ carry_out) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > A.cap_R) return;
    const RunView V = make_view(A);
    if (r == 0 && *d_M > cap) atomicOr(A.flags, (uint32_t)POST_OVF_EVENTS);
    if (r > V.R) return;
    uint32_t idx = offsets[r];
    run_events(V, r, keep_dropped != 0, [&](uint32_t pos, int v, int d, int type) {
        if (idx < cap) {
            EventRec ev;
            ev.rel_pos = pos;
            ev.d = (uint16_t)d;
            ev.v = (int8_t)v;
            ev.type = (int8_t)type;
            out[idx] = ev;
        }
        idx++;
    });
    if (r == V.R && carry_out) {  // (cur_state, last_bit, dur) at w1
        RunCarry c;
        if (V.w1 > V.w0) {
            c.st = V.s_end(r);
            c.last_bit = V.u(r);
            c.dur = (int)((uint32_t)(V.w1 - 1 - V.p(r)) % (uint32_t)V.mx) + 1;
        } else {
            c.st = V.st0; c.last_bit = V.lb0; c.dur = V.dur0;
        }
        c.pad = 0;
        *carry_out = c;
    }
}

// ---- host launchers ------------------------------------------------------------------------------
int launch_gather_transitions(const SegWork *d_works, const uint32_t *d_counts, uint32_t *d_offsets, int n_segs,
                              TransRec *d_dense, cudaStream_t stream) {
    if (n_segs <= 0) return 0;
    seg_offsets_kernel<<<1, 1024, 0, stream>>>(d_counts, d_offsets, n_segs);
    NFC_CUDA_CHECK(cudaGetLastError());
    gather_trans_kernel<<<n_segs, 256, 0, stream>>>(d_works, d_counts, d_offsets, d_dense);
    NFC_CUDA_CHECK(cudaGetLastError());
    return 0;
}

int launch_gather_pieces(const TransRec *const *d_src, const uint32_t *d_n, const uint32_t *d_off, int n_pieces,
                         TransRec *d_dense, cudaStream_t stream) {
    if (n_pieces <= 0) return 0;
    gather_pieces_kernel<<<n_pieces, 256, 0, stream>>>(d_src, d_n, d_off, d_dense);
    NFC_CUDA_CHECK(cudaGetLastError());
    return 0;
}

// Counts the events of every run into d_counts[0..cap_R] and scans them into d_offsets; *d_total = number of events.
// d_R, d_rc_in: device memory (see RunArgs).
int launch_run_count(const TransRec *d_tr, const uint32_t *d_R, uint32_t cap_R, int64_t w0, int64_t w1, const RunCarry *d_rc_in,
                     int mx, int keep_dropped, uint32_t *d_counts, uint32_t *d_offsets, uint32_t *d_scratch, uint32_t *d_total,
                     uint32_t *d_flags, cudaStream_t stream) {
    RunArgs A;
    A.tr = d_tr; A.d_R = d_R; A.cap_R = cap_R; A.w0 = (int32_t)w0; A.w1 = (int32_t)w1; A.rc_in = d_rc_in; A.mx = mx; A.flags = d_flags;
    const uint32_t nrun = cap_R + 1;
    run_count_kernel<<<(nrun + 255) / 256, 256, 0, stream>>>(A, keep_dropped, d_counts);
    NFC_CUDA_CHECK(cudaGetLastError());
    return device_exclusive_scan<uint32_t, AddU32>(d_counts, d_offsets, nrun, 0u, AddU32(), d_scratch, d_total, stream);
}

int launch_run_write(const TransRec *d_tr, const uint32_t *d_R, uint32_t cap_R, int64_t w0, int64_t w1, const RunCarry *d_rc_in,
                     int mx, int keep_dropped, const uint32_t *d_offsets, EventRec *d_events, uint32_t cap, const uint32_t *d_M,
                     RunCarry *d_carry_out, uint32_t *d_flags, cudaStream_t stream) {
    RunArgs A;
    A.tr = d_tr; A.d_R = d_R; A.cap_R = cap_R; A.w0 = (int32_t)w0; A.w1 = (int32_t)w1; A.rc_in = d_rc_in; A.mx = mx; A.flags = d_flags;
    const uint32_t nrun = cap_R + 1;
    run_write_kernel<<<(nrun + 255) / 256, 256, 0, stream>>>(A, keep_dropped, d_offsets, d_events, cap, d_M, d_carry_out);
    NFC_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace nfc
