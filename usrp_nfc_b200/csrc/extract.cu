// extract.cu -- class bitmap of a slab -> dense, ordered transition list.
//
// The streaming slicer (slicer_fast.cuh) writes two bits per sample: val != -1 and val == 1, eight words per
// chunk of 128 samples (sample 4l+j of the chunk at bit l of word j / word 4+j).  A transition is a sample
// whose val differs from its predecessor's (transition_sink.py:84-92: `val != last_bit`); the first sample of
// the slab is compared with the carried `_last_bit`.  One thread per chunk: count -> scan of block totals ->
// write, in stream order.
#include "common.cuh"
#include "scan.cuh"

namespace nfc {

static const int EX_BLOCK = 256;

struct ChunkMaps {
    uint32_t T[4];   // transition at sample 4l+j <-> bit l of T[j], already restricted to [a, b)
    uint4 nl, h;
};

__device__ __forceinline__ uint32_t lane_range_mask(int64_t lo, int64_t hi) {  // bits [lo, hi) of a word, clamped
    if (lo < 0) lo = 0;
    if (hi > 32) hi = 32;
    if (hi <= lo) return 0u;
    const uint32_t upto_hi = hi >= 32 ? 0xffffffffu : ((1u << (int)hi) - 1u);
    const uint32_t below_lo = lo >= 32 ? 0xffffffffu : ((1u << (int)lo) - 1u);
    return upto_hi & ~below_lo;
}

__device__ __forceinline__ int class_at(const uint4 &nl, const uint4 &h, int l, int j) {
    const uint32_t wn = j == 0 ? nl.x : (j == 1 ? nl.y : (j == 2 ? nl.z : nl.w));
    const uint32_t wh = j == 0 ? h.x : (j == 1 ? h.y : (j == 2 ? h.z : h.w));
    return (int)((wn >> l) & 1u) + (int)((wh >> l) & 1u) - 1;
}

// k: chunk index in the bitmap; cpos: stream position of its first sample; [a, b): the slab
__device__ __forceinline__ void load_chunk(const uint32_t *__restrict__ bm, int64_t k, int64_t cpos, int64_t a, int64_t b,
                                           int carry_val, ChunkMaps &m) {
    const uint4 *p = reinterpret_cast<const uint4 *>(bm + k * 8);
    m.nl = __ldg(p);
    m.h = __ldg(p + 1);
    uint32_t pnl = 0u, ph = 0u;
    if (cpos > a) {  // the sample before this chunk is in the slab: last sample of the previous chunk
        pnl = __ldg(bm + (k - 1) * 8 + 3) >> 31;
        ph = __ldg(bm + (k - 1) * 8 + 7) >> 31;
    }
    m.T[0] = (m.nl.x ^ ((m.nl.w << 1) | pnl)) | (m.h.x ^ ((m.h.w << 1) | ph));
    m.T[1] = (m.nl.y ^ m.nl.x) | (m.h.y ^ m.h.x);
    m.T[2] = (m.nl.z ^ m.nl.y) | (m.h.z ^ m.h.y);
    m.T[3] = (m.nl.w ^ m.nl.z) | (m.h.w ^ m.h.z);
    if (cpos < a || cpos + 128 > b) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            // sample 4l + j is inside [a, b)  <=>  l in [ceil((a - cpos - j) / 4), ceil((b - cpos - j) / 4))
            const int64_t lo = (a - cpos - j + 3) >> 2, hi = (b - cpos - j + 3) >> 2;
            m.T[j] &= lane_range_mask(lo, hi);
        }
    }
    if (a >= cpos && a < cpos + 128 && a < b) {  // the slab's first sample: compared with the carried val
        const int off = (int)(a - cpos), l = off >> 2, j = off & 3;
        const bool tr = class_at(m.nl, m.h, l, j) != carry_val;
#pragma unroll
        for (int jj = 0; jj < 4; jj++)
            if (jj == j) m.T[jj] = (m.T[jj] & ~(1u << l)) | ((tr ? 1u : 0u) << l);
    }
}

// val of the sample before the slab: the carry the slab before left on the device (its `_last_bit`), or -- when that slab's
// run kernels may not have finished yet and the sample is covered by the same bitmap -- the val bits of sample a - 1
__device__ __forceinline__ int carry_in(const uint32_t *__restrict__ bm, int64_t bm_pos0, int64_t a, const RunCarry *__restrict__ rc_in,
                                        int carry_from_bm) {
    if (!carry_from_bm) return rc_in->last_bit;
    const int64_t q = a - 1 - bm_pos0;
    const int64_t k = q >> 7;
    const int off = (int)(q & 127), l = off >> 2, jj = off & 3;
    return (int)((__ldg(bm + k * 8 + jj) >> l) & 1u) + (int)((__ldg(bm + k * 8 + 4 + jj) >> l) & 1u) - 1;
}

// block_counts[blockIdx] = transitions in this block's chunks
__global__ void __launch_bounds__(EX_BLOCK) extract_count_kernel(const uint32_t *__restrict__ bm, int64_t k0, int64_t nchunks,
                                                                 int64_t bm_pos0, int64_t a, int64_t b,
                                                                 const RunCarry *__restrict__ rc_in, int carry_from_bm,
                                                                 uint32_t *__restrict__ block_counts) {
    const int64_t i = (int64_t)blockIdx.x * EX_BLOCK + threadIdx.x;
    const int carry_val = carry_in(bm, bm_pos0, a, rc_in, carry_from_bm);
    int cnt = 0;
    if (i < nchunks) {
        ChunkMaps m;
        const int64_t k = k0 + i;
        load_chunk(bm, k, bm_pos0 + k * 128, a, b, carry_val, m);
        cnt = __popc(m.T[0]) + __popc(m.T[1]) + __popc(m.T[2]) + __popc(m.T[3]);
    }
    __shared__ int wsum[EX_BLOCK / 32];
    const int s = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
#pragma unroll
        for (int q = 0; q < EX_BLOCK / 32; q++) t += wsum[q];
        block_counts[blockIdx.x] = (uint32_t)t;
    }
}

__global__ void __launch_bounds__(EX_BLOCK) extract_write_kernel(const uint32_t *__restrict__ bm, int64_t k0, int64_t nchunks,
                                                                 int64_t bm_pos0, int64_t a, int64_t b,
                                                                 const RunCarry *__restrict__ rc_in, int carry_from_bm,
                                                                 const uint32_t *__restrict__ block_offsets,
                                                                 TransRec *__restrict__ out, uint32_t out_cap) {
    const int64_t i = (int64_t)blockIdx.x * EX_BLOCK + threadIdx.x;
    const int carry_val = carry_in(bm, bm_pos0, a, rc_in, carry_from_bm);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    ChunkMaps m;
    m.T[0] = m.T[1] = m.T[2] = m.T[3] = 0u;
    int64_t cpos = 0;
    if (i < nchunks) {
        const int64_t k = k0 + i;
        cpos = bm_pos0 + k * 128;
        load_chunk(bm, k, cpos, a, b, carry_val, m);
    }
    const int cnt = __popc(m.T[0]) + __popc(m.T[1]) + __popc(m.T[2]) + __popc(m.T[3]);
    // exclusive scan of cnt over the block
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    __shared__ int wsum[EX_BLOCK / 32];
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    int wbase = 0;
#pragma unroll
    for (int q = 0; q < EX_BLOCK / 32; q++)
        if (q < warp) wbase += wsum[q];
    uint32_t idx = block_offsets[blockIdx.x] + (uint32_t)(wbase + inc - cnt);
    if (cnt == 0) return;
    const uint32_t rel0 = (uint32_t)(cpos - a);  // wraps for the chunk holding `a`; rel0 + m is right modulo 2^32
    uint32_t T0 = m.T[0], T1 = m.T[1], T2 = m.T[2], T3 = m.T[3];
    while (T0 | T1 | T2 | T3) {
        const int p0 = T0 ? ((__ffs(T0) - 1) << 2) : 1000, p1 = T1 ? (((__ffs(T1) - 1) << 2) | 1) : 1000;
        const int p2 = T2 ? (((__ffs(T2) - 1) << 2) | 2) : 1000, p3 = T3 ? (((__ffs(T3) - 1) << 2) | 3) : 1000;
        const int pm = min(min(p0, p1), min(p2, p3));
        const int l = pm >> 2, j = pm & 3;
        const uint32_t clr = ~(1u << l);
        if (j == 0) T0 &= clr; else if (j == 1) T1 &= clr; else if (j == 2) T2 &= clr; else T3 &= clr;
        if (idx < out_cap) out[idx] = pack_trans(rel0 + (uint32_t)pm, class_at(m.nl, m.h, l, j));
        idx++;
    }
}

// Counts the transitions of [a, b) (device total in *d_total) and leaves per-block offsets in d_block_offsets.
// d_rc_in: the carry the slab before left on the device (its last_bit is compared with the slab's first sample).
int launch_extract_count(const uint32_t *d_bm, int64_t bm_pos0, int64_t a, int64_t b, const RunCarry *d_rc_in, int carry_from_bm,
                         uint32_t *d_block_counts,
                         uint32_t *d_block_offsets, uint32_t *d_scan_scratch, uint32_t *d_total, cudaStream_t stream) {
    if (b <= a) return 0;
    const int64_t k0 = (a - bm_pos0) >> 7, k1 = (b - 1 - bm_pos0) >> 7;
    const int64_t nchunks = k1 - k0 + 1;
    const unsigned nblk = (unsigned)((nchunks + EX_BLOCK - 1) / EX_BLOCK);
    extract_count_kernel<<<nblk, EX_BLOCK, 0, stream>>>(d_bm, k0, nchunks, bm_pos0, a, b, d_rc_in, carry_from_bm, d_block_counts);
    NFC_CUDA_CHECK(cudaGetLastError());
    return device_exclusive_scan<uint32_t, AddU32>(d_block_counts, d_block_offsets, nblk, 0u, AddU32(), d_scan_scratch, d_total, stream);
}

int launch_extract_write(const uint32_t *d_bm, int64_t bm_pos0, int64_t a, int64_t b, const RunCarry *d_rc_in, int carry_from_bm,
                         const uint32_t *d_block_offsets, TransRec *d_out, uint32_t out_cap, cudaStream_t stream) {
    if (b <= a) return 0;
    const int64_t k0 = (a - bm_pos0) >> 7, k1 = (b - 1 - bm_pos0) >> 7;
    const int64_t nchunks = k1 - k0 + 1;
    const unsigned nblk = (unsigned)((nchunks + EX_BLOCK - 1) / EX_BLOCK);
    extract_write_kernel<<<nblk, EX_BLOCK, 0, stream>>>(d_bm, k0, nchunks, bm_pos0, a, b, d_rc_in, carry_from_bm, d_block_offsets, d_out, out_cap);
    NFC_CUDA_CHECK(cudaGetLastError());
    return 0;
}

size_t extract_blocks(int64_t bm_pos0, int64_t a, int64_t b) {
    if (b <= a) return 0;
    const int64_t k0 = (a - bm_pos0) >> 7, k1 = (b - 1 - bm_pos0) >> 7;
    return (size_t)((k1 - k0 + 1 + EX_BLOCK - 1) / EX_BLOCK);
}

}  // namespace nfc
