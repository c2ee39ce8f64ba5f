// synth.cu -- device-side rendering of synthetic ISO 14443A traffic.
//
// Stands in for binary_src.work (binary_src.py:64-103): a pulse schedule (level code, length in
// samples) produced on the host from the line-code encoders is expanded to samples, with
// multiplicative Gaussian noise and a slow fade, quantised to int16 and normalised the way a 16-bit
// WAV recording would be delivered.  Used to build captures too large to synthesise on the host
// (1e10 samples); the schedule repeats, the noise does not (it is a hash of the sample index).
#include "common.cuh"

namespace nfc {

__device__ __forceinline__ uint32_t mix32(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return (uint32_t)x;
}

__global__ void synth_kernel(float *__restrict__ out, int64_t n, int64_t first, const int8_t *__restrict__ codes,
                             const int64_t *__restrict__ ends, int64_t n_runs, int64_t period, float carrier, float m_pause,
                             float m_high, float noise, float fade, double fade_period, uint64_t seed, int as_envelope) {
    const int PER = 16;
    const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * PER;
    if (i0 >= n) return;
    int64_t j = (first + i0) % period;
    // first run whose end is > j
    int64_t lo = 0, hi = n_runs - 1;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (ends[mid] > j) hi = mid; else lo = mid + 1;
    }
    int64_t run = lo;
    int64_t run_end = ends[run];
    int code = codes[run];
    for (int k = 0; k < PER && i0 + k < n; k++) {
        const int64_t i = i0 + k;
        const int64_t gi = first + i;  // index in the endless capture
        while (j >= run_end) {
            run++;
            if (run >= n_runs) { run = 0; j -= period; }
            run_end = ends[run];
            code = codes[run];
        }
        float amp = carrier * (code == 0 ? 1.0f : (code == 1 ? m_pause : m_high));
        if (noise != 0.0f) {
            const uint32_t a = mix32((uint64_t)gi * 2 + seed), b = mix32((uint64_t)gi * 2 + 1 + seed * 0x9e3779b97f4a7c15ULL);
            const float u1 = ((float)a + 1.0f) * 2.3283064e-10f, u2 = (float)b * 2.3283064e-10f;
            const float g = sqrtf(-2.0f * __logf(u1)) * __cosf(6.2831853f * u2);
            amp *= 1.0f + noise * g;
        }
        if (fade != 0.0f) amp *= 1.0f + fade * __sinf((float)(6.283185307179586 * (double)(gi % (int64_t)fade_period) / fade_period));
        float q = rintf(amp * 32767.0f);
        q = fminf(fmaxf(q, -32768.0f), 32767.0f);
        const float x = __fdiv_rn(q, 32767.0f);
        out[i] = as_envelope ? __fmul_rn(x, x) : x;
        j++;
    }
}

int synth_render(void *dev_out, int64_t n, int64_t first_index, const int8_t *codes, const int64_t *lens, int64_t n_runs, float carrier,
                 float pause, float tag_high, float noise, float fade, double fade_period, uint64_t seed,
                 int as_envelope, cudaStream_t stream) {
    if (n <= 0 || n_runs <= 0) return 0;
    int64_t *h_ends = (int64_t *)malloc(sizeof(int64_t) * (size_t)n_runs);
    if (!h_ends) { set_error("out of memory"); return -1; }
    int64_t acc = 0;
    for (int64_t r = 0; r < n_runs; r++) { acc += lens[r]; h_ends[r] = acc; }
    int8_t *d_codes = nullptr;
    int64_t *d_ends = nullptr;
    cudaError_t e1 = cudaMalloc(&d_codes, (size_t)n_runs), e2 = cudaMalloc(&d_ends, sizeof(int64_t) * (size_t)n_runs);
    if (e1 != cudaSuccess || e2 != cudaSuccess || acc <= 0) {
        free(h_ends);
        if (d_codes) cudaFree(d_codes);
        if (d_ends) cudaFree(d_ends);
        set_error("synth_render: allocation failed or empty schedule");
        return -1;
    }
    cudaMemcpy(d_codes, codes, (size_t)n_runs, cudaMemcpyHostToDevice);
    cudaMemcpy(d_ends, h_ends, sizeof(int64_t) * (size_t)n_runs, cudaMemcpyHostToDevice);
    free(h_ends);
    const int64_t threads = (n + 15) / 16;
    const unsigned blocks = (unsigned)((threads + 255) / 256);
    synth_kernel<<<blocks, 256, 0, stream>>>((float *)dev_out, n, first_index, d_codes, d_ends, n_runs, acc, carrier, pause / carrier,
                                             tag_high, noise, fade, fade_period > 1 ? fade_period : 1.0, seed, as_envelope);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(d_codes);
    cudaFree(d_ends);
    if (e != cudaSuccess) { set_error("synth kernel failed: %s", cudaGetErrorString(e)); return -1; }
    return 0;
}

}  // namespace nfc
