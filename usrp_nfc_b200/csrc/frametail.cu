// frametail.cu -- the per-frame tail of the decode path on the device: what fsm.process_bits does to a frame before
// any protocol logic (fsm.py:51-66 _fix_ending, :28-49 _check_parity, :114-131 _print_enc) plus CRC_A
// (utilities.py:26-46).  One thread per frame (frames are at most a few hundred bits), three launches: lengths ->
// exclusive scan of the byte counts -> bytes, parity flags, verdicts.  Crypto1 traffic still has to be decrypted on
// the host before its parity means anything; for plain traffic the verdicts are final.
#include "../../include/usrp_nfc_b200.h"
#include "common.cuh"
#include "scan.cuh"

namespace nfc {

static const int FT_BLOCK = 128;

struct FrameIn {  // nfc_frame
    int64_t pos;
    int64_t bit_off;
    int32_t nbits;
    int32_t type;
};
static_assert(sizeof(FrameIn) == sizeof(nfc_frame), "layout of nfc_frame");

__device__ __forceinline__ int start_bit_of(int type) { return type == 0 ? 1 : 0; }  // packets.py:24-30: TAG -> 1, READER -> 0

// tails[i].nbits / nbytes / fix_flag and counts[i] = nbytes
__global__ void __launch_bounds__(FT_BLOCK) frametail_len_kernel(const FrameIn *__restrict__ frames, int64_t n,
                                                                const uint8_t *__restrict__ bits_tag,
                                                                const uint8_t *__restrict__ bits_reader,
                                                                nfc_frame_tail *__restrict__ tails,
                                                                uint32_t *__restrict__ counts) {
    const int64_t i = (int64_t)blockIdx.x * FT_BLOCK + threadIdx.x;
    if (i >= n) return;
    const FrameIn f = frames[i];
    const int ll = f.nbits, rem = ll % 9;
    int fixed = ll, flag = 0;
    if (rem == 8) {
        fixed = ll + 1;  // "missing final one -- assume same as end bit"
    } else if (rem == 1) {
        const uint8_t *b = (f.type == 0 ? bits_tag : bits_reader) + f.bit_off;
        if ((int)b[ll - 1] != start_bit_of(f.type)) flag = 1;  // "EXTRA ERROR"
        fixed = ll - 1;
    } else if (rem != 0) {
        flag = 2;  // "MANY MORE ERROR"
        fixed = ll - rem;
    }
    nfc_frame_tail t;
    t.nbits = fixed;
    t.nbytes = fixed / 9;
    t.byte_off = 0;
    t.fix_flag = (int8_t)flag;
    t.parity_ok = 0;
    t.crc_ok = 0;
    for (int k = 0; k < 5; k++) t.pad[k] = 0;
    tails[i] = t;
    counts[i] = (uint32_t)t.nbytes;
}

__global__ void __launch_bounds__(FT_BLOCK) frametail_write_kernel(const FrameIn *__restrict__ frames, int64_t n,
                                                                  const uint8_t *__restrict__ bits_tag,
                                                                  const uint8_t *__restrict__ bits_reader,
                                                                  const uint32_t *__restrict__ offsets,
                                                                  nfc_frame_tail *__restrict__ tails, uint8_t *__restrict__ bytes,
                                                                  uint8_t *__restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * FT_BLOCK + threadIdx.x;
    if (i >= n) return;
    const FrameIn f = frames[i];
    const uint8_t *b = (f.type == 0 ? bits_tag : bits_reader) + f.bit_off;
    const int nbytes = tails[i].nbytes, ll = f.nbits;
    const uint32_t off = offsets[i];
    const int sb = start_bit_of(f.type);
    bool all_ok = nbytes > 0;  // fsm.process_bits treats an empty list like None (fsm.py:226-228)
    uint32_t wcrc = 0x6363u;   // CRC_14443_A over all bytes but the last two
    uint32_t last2[2] = {0u, 0u};
    for (int k = 0; k < nbytes; k++) {
        uint32_t cur = 0u, ones = 0u;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint32_t bit = b[k * 9 + j] & 1u;
            cur |= bit << j;
            ones += bit;
        }
        const int pi = k * 9 + 8;
        const uint32_t par = pi < ll ? (uint32_t)(b[pi] & 1u) : (uint32_t)sb;  // the bit _fix_ending appended
        const bool viol = (ones & 1u) == par;                                   // "if set_bits & 1 == bit" -> flagged '!'
        all_ok = all_ok && !viol;
        bytes[off + k] = (uint8_t)cur;
        flags[off + k] = viol ? 1 : 0;
        if (k + 2 < nbytes) {
            uint32_t x = cur ^ (wcrc & 0xffu);
            x = x ^ ((x << 4) & 0xffu);
            wcrc = (wcrc >> 8) ^ (x << 8) ^ (x << 3) ^ (x >> 4);
        } else {
            last2[k + 2 - nbytes] = cur;
        }
    }
    nfc_frame_tail t = tails[i];
    t.byte_off = (int64_t)off;
    t.parity_ok = all_ok ? 1 : 0;
    t.crc_ok = (nbytes >= 2 && (wcrc & 0xffu) == last2[0] && ((wcrc >> 8) & 0xffu) == last2[1]) ? 1 : 0;
    tails[i] = t;
}

struct DevMem {
    void *p = nullptr;
    ~DevMem() { if (p) cudaFree(p); }
    int alloc(size_t bytes) { NFC_CUDA_CHECK(cudaMalloc(&p, bytes ? bytes : 16)); return 0; }
    template <class T> T *as() { return static_cast<T *>(p); }
};

static int64_t frames_tail(int device, const nfc_frame *frames, int64_t n, const uint8_t *bits_tag, int64_t n_bits_tag,
                           const uint8_t *bits_reader, int64_t n_bits_reader, nfc_frame_tail *tails, uint8_t *bytes,
                           uint8_t *parity_flags, int64_t bytes_cap) {
    if (n < 0 || n_bits_tag < 0 || n_bits_reader < 0 || (n > 0 && (!frames || !tails))) {
        set_error("frames_tail: bad arguments");
        return -1;
    }
    if (n == 0) return 0;
    // the records are checked on the host before anything is read through them on the device
    int64_t max_bytes = 0;
    for (int64_t i = 0; i < n; i++) {
        const nfc_frame &f = frames[i];
        const int64_t lim = f.type == 0 ? n_bits_tag : n_bits_reader;
        if ((f.type != 0 && f.type != 1) || f.nbits < 0 || f.bit_off < 0 || f.bit_off + f.nbits > lim ||
            (f.nbits > 0 && !(f.type == 0 ? bits_tag : bits_reader))) {
            set_error("frames_tail: frame %lld does not lie inside its bit buffer", (long long)i);
            return -1;
        }
        max_bytes += (f.nbits + 1) / 9;
    }
    if (max_bytes > (int64_t)0xffffffffll) {
        set_error("frames_tail: more than 2^32 bytes in one call");
        return -1;
    }
    NFC_CUDA_CHECK(cudaSetDevice(device));
    cudaStream_t st;
    NFC_CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    struct StreamGuard {
        cudaStream_t s;
        ~StreamGuard() { cudaStreamDestroy(s); }
    } guard{st};
    DevMem d_fr, d_bt, d_br, d_tails, d_cnt, d_off, d_scr, d_tot, d_bytes, d_flags;
    const size_t scr = scan_scratch_elems((size_t)n) + 16;
    if (d_fr.alloc(sizeof(nfc_frame) * (size_t)n) || d_bt.alloc((size_t)n_bits_tag) || d_br.alloc((size_t)n_bits_reader) ||
        d_tails.alloc(sizeof(nfc_frame_tail) * (size_t)n) || d_cnt.alloc(sizeof(uint32_t) * (size_t)n) ||
        d_off.alloc(sizeof(uint32_t) * (size_t)n) || d_scr.alloc(sizeof(uint32_t) * scr) || d_tot.alloc(sizeof(uint32_t)))
        return -1;
    NFC_CUDA_CHECK(cudaMemcpyAsync(d_fr.p, frames, sizeof(nfc_frame) * (size_t)n, cudaMemcpyHostToDevice, st));
    if (n_bits_tag) NFC_CUDA_CHECK(cudaMemcpyAsync(d_bt.p, bits_tag, (size_t)n_bits_tag, cudaMemcpyHostToDevice, st));
    if (n_bits_reader) NFC_CUDA_CHECK(cudaMemcpyAsync(d_br.p, bits_reader, (size_t)n_bits_reader, cudaMemcpyHostToDevice, st));
    const unsigned nblk = (unsigned)((n + FT_BLOCK - 1) / FT_BLOCK);
    frametail_len_kernel<<<nblk, FT_BLOCK, 0, st>>>(d_fr.as<FrameIn>(), n, d_bt.as<uint8_t>(), d_br.as<uint8_t>(),
                                                   d_tails.as<nfc_frame_tail>(), d_cnt.as<uint32_t>());
    NFC_CUDA_CHECK(cudaGetLastError());
    if (device_exclusive_scan<uint32_t, AddU32>(d_cnt.as<uint32_t>(), d_off.as<uint32_t>(), (size_t)n, 0u, AddU32(),
                                                d_scr.as<uint32_t>(), d_tot.as<uint32_t>(), st))
        return -1;
    uint32_t total = 0;
    NFC_CUDA_CHECK(cudaMemcpyAsync(&total, d_tot.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    NFC_CUDA_CHECK(cudaStreamSynchronize(st));
    if ((int64_t)total > bytes_cap || (total > 0 && (!bytes || !parity_flags))) {
        set_error("frames_tail: byte buffers too small (%lld needed, %lld given)", (long long)total, (long long)bytes_cap);
        return -1;
    }
    if (d_bytes.alloc(total) || d_flags.alloc(total)) return -1;
    frametail_write_kernel<<<nblk, FT_BLOCK, 0, st>>>(d_fr.as<FrameIn>(), n, d_bt.as<uint8_t>(), d_br.as<uint8_t>(),
                                                     d_off.as<uint32_t>(), d_tails.as<nfc_frame_tail>(), d_bytes.as<uint8_t>(),
                                                     d_flags.as<uint8_t>());
    NFC_CUDA_CHECK(cudaGetLastError());
    NFC_CUDA_CHECK(cudaMemcpyAsync(tails, d_tails.p, sizeof(nfc_frame_tail) * (size_t)n, cudaMemcpyDeviceToHost, st));
    if (total) {
        NFC_CUDA_CHECK(cudaMemcpyAsync(bytes, d_bytes.p, total, cudaMemcpyDeviceToHost, st));
        NFC_CUDA_CHECK(cudaMemcpyAsync(parity_flags, d_flags.p, total, cudaMemcpyDeviceToHost, st));
    }
    NFC_CUDA_CHECK(cudaStreamSynchronize(st));
    return (int64_t)total;
}

}  // namespace nfc

extern "C" int64_t nfc_frames_tail(int device, const nfc_frame *frames, int64_t n_frames, const uint8_t *bits_tag,
                                   int64_t n_bits_tag, const uint8_t *bits_reader, int64_t n_bits_reader, nfc_frame_tail *tails,
                                   uint8_t *bytes, uint8_t *parity_flags, int64_t bytes_cap) {
    return nfc::frames_tail(device, frames, n_frames, bits_tag, n_bits_tag, bits_reader, n_bits_reader, tails, bytes,
                            parity_flags, bytes_cap);
}
