// scan.cuh -- order-preserving exclusive scan over an arbitrary associative operator (device-wide).
//
// Used for the small arrays behind the sample-rate kernel: event counts per run, chunk
// transfer functions of the line-code automata, output counters.  Three phases per level
// (block scan, scan of block totals, add), recursing while more than one block remains.
#pragma once

#include <cuda_runtime.h>

#include <vector>

#include "common.cuh"

namespace nfc {

static const int SCAN_BLOCK = 256;
static const int SCAN_ITEMS = 4;  // per thread
static const int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;

// generic shuffle for trivially copyable T (sizeof multiple of 4)
template <class T>
__device__ __forceinline__ T shfl_up_any(const T &v, int delta) {
    static_assert(sizeof(T) % 4 == 0, "T must be a multiple of 4 bytes");
    T r;
    const unsigned *src = reinterpret_cast<const unsigned *>(&v);
    unsigned *dst = reinterpret_cast<unsigned *>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 4); i++) dst[i] = __shfl_up_sync(0xffffffffu, src[i], delta);
    return r;
}

// out[i] = op(in[0..i-1]) within the block (identity for the first), totals[block] = op(all)
template <class T, class Op>
__global__ void scan_block_kernel(const T *__restrict__ in, T *__restrict__ out, T *__restrict__ totals, size_t n,
                                  T identity, Op op) {
    __shared__ T wtot[SCAN_BLOCK / 32];
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    T v[SCAN_ITEMS];
    T run = identity;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        v[j] = (base + j < n) ? in[base + j] : identity;
        run = op(run, v[j]);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T nb = shfl_up_any(inc, o);
        if (lane >= o) inc = op(nb, inc);
    }
    if (lane == 31) wtot[warp] = inc;
    T exc = shfl_up_any(inc, 1);
    if (lane == 0) exc = identity;
    __syncthreads();
    T wbase = identity;
    T tot = identity;
    for (int w = 0; w < SCAN_BLOCK / 32; w++) {
        if (w < warp) wbase = op(wbase, wtot[w]);
        tot = op(tot, wtot[w]);
    }
    T acc = op(wbase, exc);
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        if (base + j < n) out[base + j] = acc;
        acc = op(acc, v[j]);
    }
    if (threadIdx.x == 0 && totals) totals[blockIdx.x] = tot;
}

template <class T, class Op>
__global__ void scan_add_kernel(T *__restrict__ out, const T *__restrict__ block_excl, size_t n, Op op) {
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    const T b = block_excl[blockIdx.x];
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++)
        if (base + j < n) out[base + j] = op(b, out[base + j]);
}

static inline size_t scan_blocks(size_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE; }

// scratch elements needed for a scan of n items
static inline size_t scan_scratch_elems(size_t n) {
    size_t total = 0;
    while (n > 1) {
        size_t b = scan_blocks(n);
        total += 2 * b;
        if (b <= 1) break;
        n = b;
    }
    return total + 4;
}

// Exclusive scan of in[0..n) into out[0..n); *total_out (device, optional) = op over everything.
template <class T, class Op>
int device_exclusive_scan(const T *in, T *out, size_t n, T identity, Op op, T *scratch, T *total_out,
                          cudaStream_t stream) {
    if (n == 0) {
        if (total_out) NFC_CUDA_CHECK(cudaMemcpyAsync(total_out, &identity, 0, cudaMemcpyHostToDevice, stream));
        return 0;
    }
    const size_t nb = scan_blocks(n);
    T *totals = scratch;
    T *totals_excl = scratch + nb;
    scan_block_kernel<T, Op><<<(unsigned)nb, SCAN_BLOCK, 0, stream>>>(in, out, totals, n, identity, op);
    NFC_CUDA_CHECK(cudaGetLastError());
    if (nb == 1) {
        if (total_out) NFC_CUDA_CHECK(cudaMemcpyAsync(total_out, totals, sizeof(T), cudaMemcpyDeviceToDevice, stream));
        return 0;
    }
    if (device_exclusive_scan<T, Op>(totals, totals_excl, nb, identity, op, scratch + 2 * nb, total_out, stream)) return -1;
    scan_add_kernel<T, Op><<<(unsigned)nb, SCAN_BLOCK, 0, stream>>>(out, totals_excl, n, op);
    NFC_CUDA_CHECK(cudaGetLastError());
    return 0;
}

struct AddU32 {
    __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a + b; }
};

}  // namespace nfc
