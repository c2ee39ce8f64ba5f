/*
 * oracle/nfc_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, IEEE double, strictly sequential) of the sample-rate
 * decode path of giech/usrp_nfc.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.  The
 * product (usrp_nfc_b200/csrc) never links, imports or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement
 * against tests/golden/, which oracle/gen_golden.py produced by running the
 * reference's own transition_sink.py / miller.py / manchester.py / packets.py /
 * fsm.py (imported from /root/reference/code) and against the reference's
 * outputs/1k_with_enc.out, outputs/ultralight.out and outputs/ultralight_bits.txt.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference checkout's code/ directory).
 */
#ifndef NFC_ORACLE_H
#define NFC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* One element of the list transition_sink hands to its callback
 * (transition_sink.py:89-90,97): ((v, d*factor), cur_state-1).  The duration is
 * kept as the integer sample count d; dur_us = d * factor. */
typedef struct {
    int64_t pos;  /* 0-based stream index of the sample whose processing emitted it */
    int32_t d;    /* run length in samples, 1..max_len */
    int8_t v;     /* -1..2 */
    int8_t type;  /* -1 (dropped by background.py:30-35), 0 tag->reader, 1 reader->tag */
    int16_t pad;
} nfc_event;

/* One cpp.append_bit(value, packet_type) call (manchester.py:28, miller.py:151). */
typedef struct {
    int64_t pos; /* pos of the event that produced it */
    int8_t type; /* PacketType: 0 TAG_TO_READER, 1 READER_TO_TAG (packets.py:19-20) */
    int8_t val;  /* 0/1 data bit, or ErrorCode 2..7 (utilities.py:7-14) */
    int16_t pad;
    int32_t pad2;
} nfc_symbol;

/* One fsm.process_bits(bits, packet_type) hand-off (packets.py:94-98). */
typedef struct {
    int64_t pos;     /* pos of the event whose error symbol closed the frame */
    int64_t bit_off; /* offset of the first bit in the bit buffer (one byte per bit) */
    int32_t nbits;
    int32_t type;
} nfc_frame;

/* ---- transition_sink (transition_sink.py:10-125) ---- */
typedef struct nfc_ts nfc_ts;
nfc_ts *nfc_ts_new(double samp_rate, double lo_val, double hi_val, int av_window, int max_len);
void nfc_ts_free(nfc_ts *);
/* One work() call.  Returns items consumed (warm-up may consume fewer than n,
 * transition_sink.py:109-125).  *called_back is 1 when the reference would have
 * invoked its callback (work_stable, :101); the events of this call are then
 * available through nfc_ts_events(). */
int64_t nfc_ts_work(nfc_ts *, const float *in, int64_t n, int *called_back);
const nfc_event *nfc_ts_events(const nfc_ts *, int64_t *count);
/* state inspection (transition_sink.py:20-34,102-106) */
void nfc_ts_get_scalars(const nfc_ts *, double *ss, int *cur_state, int *dur, int *last_bit, int *index, int *filled, int *stable);
const double *nfc_ts_ring(const nfc_ts *);
/* restore the attributes saved at transition_sink.py:102-106 (checkpoint / time-shard stitching tests) */
void nfc_ts_set_state(nfc_ts *, const double *ring, double ss, int cur_state, int dur, int last_bit, int index,
                      int64_t pos);

/* ---- background grouping + decoders + framing
 *      (background.py:30-52, manchester.py:13-61, miller.py:13-197, packets.py:57-98) ---- */
typedef struct nfc_dec nfc_dec;
nfc_dec *nfc_dec_new(int decode_reader, int decode_tag);
void nfc_dec_free(nfc_dec *);
void nfc_dec_clear_outputs(nfc_dec *);
/* Feed one callback batch (what background.append receives). */
void nfc_dec_feed(nfc_dec *, const nfc_event *ev, int64_t n, double factor);
const nfc_symbol *nfc_dec_symbols(const nfc_dec *, int64_t *count);
const nfc_frame *nfc_dec_frames(const nfc_dec *, int64_t *count);
const uint8_t *nfc_dec_bits(const nfc_dec *, int64_t *count);
/* decoder / PacketProcessor attributes as small integers:
 * miller = stage | _has_started << 2 | _prev << 3 (miller.py:14-29), manch = _prev_set | (_prev + 1) << 1
 * (manchester.py:23-25), started[t] / pending bits per PacketProcessor (packets.py:63-65). */
void nfc_dec_get_state(const nfc_dec *, int *miller, int *manch, int *started, int *npending, uint8_t *pending_bits,
                       int pending_cap);
void nfc_dec_set_state(nfc_dec *, int miller, int manch, const int *started, const int *npending,
                       const uint8_t *pending_bits);

/* ---- host tail of the path (fsm.py:28-66,114-131) ---- */
/* _fix_ending: writes the repaired frame to out (capacity n+1), returns its length.
 * *flag: 0 none, 1 "EXTRA ERROR" printed, 2 "MANY MORE ERROR" printed. */
int32_t nfc_fix_ending(const uint8_t *bits, int32_t n, int type, uint8_t *out, int *flag);
/* _check_parity: returns the byte count, or -1 where the reference returns None.
 * (fsm.process_bits treats an empty list like None: "PARITY ERROR", fsm.py:226-228.) */
int32_t nfc_check_parity(const uint8_t *bits, int32_t n, uint8_t *bytes_out);
/* utilities.CRC.calculate_crc / check_crc with CRC_14443_A (utilities.py:26-46). */
void nfc_crc_a(const uint8_t *data, int32_t n, uint8_t out[2]);
int nfc_check_crc(const uint8_t *data, int32_t n);
/* _print_enc: bytes followed by a 9th bit, and whether each is flagged '!'. Returns count. */
int32_t nfc_print_enc(const uint8_t *bits, int32_t n, uint8_t *bytes_out, uint8_t *flag_out);

/* ---- encoders used to build synthetic traffic (miller.py:200-233, manchester.py:64-79) ---- */
/* Each writes (level, dur_us) pairs; returns the pair count (capacity must be >= 2*nbits+8). */
int32_t nfc_miller_encode(const uint8_t *bits, int32_t n, int8_t *level, double *dur_us);
int32_t nfc_manchester_encode(const uint8_t *bits, int32_t n, int8_t *level, double *dur_us);

/* ---- whole chain in one call, for timing the CPU baseline ---- */
typedef struct {
    int64_t n_events, n_symbols, n_frames, n_bits;
    uint64_t digest; /* FNV-1a over events, symbols and frames */
} nfc_chain_result;
/* Runs transition_sink.work in `chunk`-item calls (honouring the consumed count),
 * feeding every callback batch to the decoders, exactly as the flowgraph does. */
void nfc_chain_run(const float *in, int64_t n, double samp_rate, double lo_val, double hi_val,
                   int av_window, int max_len, int decode_reader, int decode_tag, int64_t chunk,
                   nfc_chain_result *res);

#ifdef __cplusplus
}
#endif
#endif
