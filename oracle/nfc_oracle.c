/*
 * oracle/nfc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see nfc_oracle.h).
 *
 * Sequential CPU restatement of usrp_nfc's sample-rate decode path.  Python floats
 * are IEEE doubles, so every comparison and accumulation below is written in the
 * same order and precision as the reference's Python statements.  Compile WITHOUT
 * -ffast-math and without FMA contraction (Makefile passes -ffp-contract=off).
 *
 * Note on sum(): the reference is Python 2, whose sum() adds left to right in
 * double.  (Python >= 3.12 sum() compensates; for every input whose window sum is
 * exactly representable -- all int16-derived captures -- both give the same value.)
 */
#include "nfc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ vectors */
#define VEC(T, name)                                                        \
    typedef struct { T *p; int64_t n, cap; } name;                          \
    static void name##_push(name *v, T x) {                                 \
        if (v->n == v->cap) {                                               \
            v->cap = v->cap ? v->cap * 2 : 256;                             \
            v->p = (T *)realloc(v->p, (size_t)v->cap * sizeof(T));          \
        }                                                                   \
        v->p[v->n++] = x;                                                   \
    }
VEC(nfc_event, evvec)
VEC(nfc_symbol, symvec)
VEC(nfc_frame, frvec)
VEC(uint8_t, bytevec)

/* utilities.py:7-23 */
enum { E_NO_ERROR = 0, E_TOO_SHORT = 2, E_TOO_LONG = 3, E_ENCODING = 4, E_INTERNAL = 5, E_WRONG_DUR = 6, E_GENERAL = 7 };
static const double PL_FULL = 9.44;
static const double PL_ZERO = 3.00;
#define PL_HALF (PL_FULL / 2)
#define PL_ZERO_REM (PL_FULL - PL_ZERO)
#define PL_ONE_REM (PL_HALF - PL_ZERO)
#define PL_ONE_HALF (PL_FULL + PL_HALF)

/* packets.py:19-30 */
enum { TAG_TO_READER = 0, READER_TO_TAG = 1 };
static int start_bit(int t) { return t == TAG_TO_READER ? 1 : 0; }

/* ------------------------------------------------- transition_sink.py:10-125 */
struct nfc_ts {
    int mx, L;
    double factor, lo, hi;
    double *ar;
    int index, filled;
    double ss;
    int cur_state, dur, last_bit;
    int stable;
    int64_t pos;
    evvec ev;
};

nfc_ts *nfc_ts_new(double samp_rate, double lo_val, double hi_val, int av_window, int max_len) {
    /* transition_sink.py:12-34 */
    nfc_ts *t = (nfc_ts *)calloc(1, sizeof(nfc_ts));
    t->mx = max_len;
    t->factor = 1e6 / samp_rate;
    t->dur = 1;
    t->last_bit = 0;
    t->index = 0;
    t->filled = 0;
    t->L = av_window;
    t->ar = (double *)calloc((size_t)(av_window > 0 ? av_window : 1), sizeof(double));
    t->ss = 0;
    t->cur_state = 0;
    t->lo = lo_val;
    t->hi = hi_val;
    return t;
}

void nfc_ts_free(nfc_ts *t) {
    if (!t) return;
    free(t->ar);
    free(t->ev.p);
    free(t);
}

static int64_t ts_warmup(nfc_ts *t, const float *in, int64_t n) {
    /* transition_sink.py:109-125 */
    int64_t need = t->L - t->filled;
    int64_t can = n < need ? n : need;
    for (int64_t i = 0; i < can; i++) t->ar[t->filled + i] = (double)in[i];
    t->filled += (int)can;
    if (can == need) {
        double s = 0;
        for (int i = 0; i < t->L; i++) s += t->ar[i];
        t->ss = s;
        t->dur = t->L % t->mx;
        t->stable = 1;
    }
    t->pos += can;
    return can;
}

static void ts_emit(nfc_ts *t, int v, int d, int type) {
    nfc_event e;
    e.pos = t->pos;
    e.d = d;
    e.v = (int8_t)v;
    e.type = (int8_t)type;
    e.pad = 0;
    evvec_push(&t->ev, e);
}

static int64_t ts_stable(nfc_ts *t, const float *in, int64_t n) {
    /* transition_sink.py:37-107 */
    double *ar = t->ar;
    const int length = t->L;
    int index = t->index;
    int cur_state = t->cur_state;
    double ss = t->ss;
    const double lo = t->lo, hi = t->hi;
    int dur = t->dur;
    int last_bit = t->last_bit;
    const int mx = t->mx;

    for (int64_t i = 0; i < n; i++) {
        const double bit = (double)in[i];
        const double prev = ar[index];
        const int prev_state = cur_state;
        double ratio, cur;
        int val;

        if (ss == 0) { /* :59-63 */
            if (bit == 0) ratio = 1;
            else ratio = hi + 0.1;
        } else {
            ratio = bit * (double)length / ss; /* :65 */
        }

        if (lo > ratio) { /* :67-77 */
            val = -1;
            cur = prev;
            cur_state = 2;
        } else if (cur_state != 2 && ratio > hi) {
            val = 1;
            cur = prev;
            cur_state = 1;
        } else {
            val = 0;
            cur = bit;
        }

        ar[index] = cur; /* :80-82 */
        index = (index + 1) % length;
        ss += (cur - prev);

        if (val == last_bit) { /* :84-92 */
            dur += 1;
        } else {
            int d = prev_state == 0 ? mx : dur;
            int v = cur_state == 2 ? last_bit + 1 : last_bit;
            ts_emit(t, v, d, cur_state - 1);
            dur = 1;
            last_bit = val;
        }

        if (dur > mx) { /* :95-99 */
            int v = cur_state == 2 ? last_bit + 1 : last_bit;
            ts_emit(t, v, mx, cur_state - 1);
            dur = 1;
            cur_state = 0;
        }
        t->pos++;
    }
    t->index = index; /* :102-106 */
    t->cur_state = cur_state;
    t->ss = ss;
    t->dur = dur;
    t->last_bit = last_bit;
    return n;
}

int64_t nfc_ts_work(nfc_ts *t, const float *in, int64_t n, int *called_back) {
    t->ev.n = 0;
    if (!t->stable) {
        if (called_back) *called_back = 0;
        return ts_warmup(t, in, n);
    }
    if (called_back) *called_back = 1;
    return ts_stable(t, in, n);
}

const nfc_event *nfc_ts_events(const nfc_ts *t, int64_t *count) {
    *count = t->ev.n;
    return t->ev.p;
}

void nfc_ts_get_scalars(const nfc_ts *t, double *ss, int *cur_state, int *dur, int *last_bit, int *index, int *filled, int *stable) {
    *ss = t->ss;
    *cur_state = t->cur_state;
    *dur = t->dur;
    *last_bit = t->last_bit;
    *index = t->index;
    *filled = t->filled;
    *stable = t->stable;
}

const double *nfc_ts_ring(const nfc_ts *t) { return t->ar; }

void nfc_ts_set_state(nfc_ts *t, const double *ring, double ss, int cur_state, int dur, int last_bit, int index,
                      int64_t pos) {
    for (int i = 0; i < t->L; i++) t->ar[i] = ring[i];
    t->ss = ss;
    t->cur_state = cur_state;
    t->dur = dur;
    t->last_bit = last_bit;
    t->index = index;
    t->filled = t->L;
    t->stable = 1;
    t->pos = pos;
}

/* --------------------------------------------------------- decoders + framing */
typedef struct { /* packets.py:57-79 */
    int type, start, started;
    bytevec cur;
} pproc;

typedef struct { /* manchester.py:13-61 */
    double lo, mid, hi;
    int prev_set;
    double prev;
} manch;

typedef struct { /* miller.py:13-197 */
    double prev;
    double thres, lo, hi;
    int has_started;
    double dur_0, dur_1;
    int cur_type;
} miller;

struct nfc_dec {
    int use_reader, use_tag;
    manch mc;
    miller ml;
    pproc pp[2];
    symvec sym;
    frvec fr;
    bytevec bits;
    int64_t cur_pos; /* pos of the event being processed */
};

/* packets.py:67-79 + :94-98 */
static void cpp_append_bit(nfc_dec *d, int bit, int type) {
    nfc_symbol s;
    s.pos = d->cur_pos;
    s.type = (int8_t)type;
    s.val = (int8_t)bit;
    s.pad = 0;
    s.pad2 = 0;
    symvec_push(&d->sym, s);

    pproc *pp = &d->pp[type];
    if (bit != 0 && bit != 1) {
        if (pp->started) {
            /* cur = self._cur; self._reset_packet(); return cur */
            if (pp->cur.n > 0) { /* `if ret:` -- an empty list is not forwarded (packets.py:97) */
                nfc_frame f;
                f.pos = d->cur_pos;
                f.bit_off = d->bits.n;
                f.nbits = (int32_t)pp->cur.n;
                f.type = type;
                for (int64_t i = 0; i < pp->cur.n; i++) bytevec_push(&d->bits, pp->cur.p[i]);
                frvec_push(&d->fr, f);
            }
            pp->started = 0;
            pp->cur.n = 0;
        }
    } else {
        if (!pp->started && bit == pp->start) pp->started = 1;
        else bytevec_push(&pp->cur, (uint8_t)bit); /* "check logic" branch, packets.py:77-78 */
    }
}

/* manchester.py:23-25 */
static void manch_reset(manch *m) {
    m->prev_set = 0;
    m->prev = 0;
}

/* manchester.py:30-61, one transition */
static void manch_step(nfc_dec *d, double cur, double dur) {
    manch *m = &d->mc;
    int err = E_NO_ERROR;
    if (dur < m->lo) err = E_TOO_SHORT;
    else if (dur > m->hi) err = E_TOO_LONG;
    if (err != E_NO_ERROR) {
        manch_reset(m);
        cpp_append_bit(d, err, TAG_TO_READER);
        return;
    }
    int dual = dur > m->mid;
    double prev = m->prev;
    if (m->prev_set) {
        if (prev == cur || (prev != 0 && prev != 1)) {
            cpp_append_bit(d, E_INTERNAL, TAG_TO_READER);
            return;
        }
        cpp_append_bit(d, (int)prev, TAG_TO_READER);
        m->prev_set = dual;
    } else {
        if (dual) {
            cpp_append_bit(d, E_ENCODING, TAG_TO_READER);
            return;
        }
        m->prev_set = 1;
    }
    m->prev = cur;
}

/* miller.py:14-17 */
enum { ST_BEGINNING = 0, ST_ZERO_STAGE_0 = 1, ST_ONE_STAGE_0 = 2, ST_ONE_STAGE_1 = 3 };

static int ml_get_stage(const miller *m) { /* miller.py:31-41 */
    if (m->dur_0 == 0) return ST_BEGINNING;
    if (m->cur_type == 0) return ST_ZERO_STAGE_0;
    if (m->dur_1 == 0) return ST_ONE_STAGE_0;
    return ST_ONE_STAGE_1;
}

static void ml_set_stage(miller *m, int stage) { /* miller.py:43-59 */
    if (stage == ST_BEGINNING) {
        m->dur_0 = 0;
        m->dur_1 = 0;
        m->cur_type = 0;
    } else if (stage == ST_ZERO_STAGE_0) {
        m->dur_0 = PL_ZERO;
        m->cur_type = 0;
    } else if (stage == ST_ONE_STAGE_0) {
        m->dur_0 = PL_HALF;
        m->cur_type = 1;
    } else {
        m->dur_0 = PL_HALF;
        m->dur_1 = PL_ZERO;
        m->cur_type = 1;
    }
}

static int ml_close(const miller *m, double dur, double av) { return fabs(dur - av) <= m->thres; } /* :62-63 */

static void ml_reset(miller *m) { /* :65-67 */
    m->has_started = 0;
    ml_set_stage(m, ST_BEGINNING);
}

/* miller.py:153-197, one transition */
static void miller_step(nfc_dec *d, double cur, double dur) {
    miller *m = &d->ml;
    int rets[4];
    int nr = 0;

    if (cur == 0 && fabs(dur - PL_ZERO) < PL_ZERO / 2) dur = PL_ZERO; /* :157-158 */

    int err = E_NO_ERROR;
    int stage = ml_get_stage(m);
    if ((dur < m->lo || dur > m->hi) && (stage == ST_ZERO_STAGE_0 || stage == ST_ONE_STAGE_1)) { /* :165-167 */
        cpp_append_bit(d, m->cur_type, READER_TO_TAG);
        err = E_TOO_LONG;
    } else if (dur < m->lo) {
        err = E_TOO_SHORT;
    } else if (dur > m->hi) {
        err = E_TOO_LONG;
    }
    if (err != E_NO_ERROR) { /* :173-176 */
        cpp_append_bit(d, err, READER_TO_TAG);
        ml_reset(m);
        return;
    }

    if (stage == ST_BEGINNING) { /* handle_beginning :73-96 */
        if (cur == 0) {
            if (ml_close(m, dur, PL_ZERO)) {
                ml_set_stage(m, ST_ZERO_STAGE_0);
                m->has_started = 1;
            } else {
                rets[nr++] = E_TOO_LONG;
            }
        } else if (m->has_started) {
            int bit = 0;
            if (m->prev == 0) bit = E_ENCODING;
            if (ml_close(m, dur, PL_HALF)) {
                ml_set_stage(m, ST_ONE_STAGE_0);
            } else if (ml_close(m, dur, PL_FULL)) {
                rets[nr++] = bit;
            } else if (ml_close(m, dur, PL_ONE_HALF)) {
                rets[nr++] = bit;
                ml_set_stage(m, ST_ONE_STAGE_0);
            } else {
                rets[nr++] = E_WRONG_DUR;
            }
        }
    } else if (stage == ST_ZERO_STAGE_0) { /* handle_zs0 :98-112 */
        if (cur == 0) {
            rets[nr++] = E_ENCODING;
        } else if (ml_close(m, dur, PL_ZERO_REM)) {
            ml_set_stage(m, ST_BEGINNING);
            rets[nr++] = 0;
        } else if (ml_close(m, dur, PL_ZERO_REM + PL_HALF)) {
            ml_set_stage(m, ST_ONE_STAGE_0);
            rets[nr++] = 0;
        } else {
            rets[nr++] = E_WRONG_DUR;
        }
    } else if (stage == ST_ONE_STAGE_0) { /* handle_os0 :114-122 */
        if (cur != 0) rets[nr++] = E_ENCODING;
        else if (!ml_close(m, dur, PL_ZERO)) rets[nr++] = E_WRONG_DUR;
        else ml_set_stage(m, ST_ONE_STAGE_1);
    } else { /* handle_os1 :124-148 */
        if (cur != 1) {
            rets[nr++] = E_ENCODING;
        } else if (ml_close(m, dur, PL_ONE_REM)) {
            rets[nr++] = 1;
            ml_set_stage(m, ST_BEGINNING);
        } else {
            rets[nr++] = 1;
            ml_set_stage(m, ST_BEGINNING);
            dur -= PL_ONE_REM;
            if (ml_close(m, dur, PL_FULL)) {
                rets[nr++] = 0;
            } else if (ml_close(m, dur, PL_HALF)) {
                ml_set_stage(m, ST_ONE_STAGE_0);
            } else if (ml_close(m, dur, PL_ONE_HALF)) {
                rets[nr++] = 0;
                ml_set_stage(m, ST_ONE_STAGE_0);
            } else {
                rets[nr++] = E_WRONG_DUR;
            }
        }
    }

    for (int i = 0; i < nr; i++) { /* :191-197 */
        cpp_append_bit(d, rets[i], READER_TO_TAG);
        if (rets[i] > 1) {
            ml_reset(m);
            m->prev = 0;
        } else {
            m->prev = rets[i];
        }
    }
}

nfc_dec *nfc_dec_new(int decode_reader, int decode_tag) {
    nfc_dec *d = (nfc_dec *)calloc(1, sizeof(nfc_dec));
    d->use_reader = decode_reader;
    d->use_tag = decode_tag;
    /* manchester.py:15-21 */
    d->mc.lo = PL_HALF - 1;
    d->mc.mid = PL_HALF + 1;
    d->mc.hi = 2 * PL_HALF + 1;
    manch_reset(&d->mc);
    /* miller.py:19-29 */
    d->ml.prev = 0;
    d->ml.thres = 1.5;
    d->ml.lo = PL_ZERO - d->ml.thres;
    d->ml.hi = 2 * PL_FULL;
    d->ml.dur_1 = 0;
    ml_reset(&d->ml);
    /* packets.py:57-66,81-92 */
    for (int i = 0; i < 2; i++) {
        d->pp[i].type = i;
        d->pp[i].start = start_bit(i);
        d->pp[i].started = 0;
    }
    return d;
}

void nfc_dec_free(nfc_dec *d) {
    if (!d) return;
    free(d->pp[0].cur.p);
    free(d->pp[1].cur.p);
    free(d->sym.p);
    free(d->fr.p);
    free(d->bits.p);
    free(d);
}

void nfc_dec_clear_outputs(nfc_dec *d) {
    d->sym.n = 0;
    d->fr.n = 0;
    d->bits.n = 0;
}

/* background.py:30-35 applied to one event of a same-type group */
static void dec_one(nfc_dec *d, const nfc_event *e, double factor, int group_type) {
    double dur = (double)e->d * factor; /* transition_sink.py:89,97: d*factor */
    d->cur_pos = e->pos;
    if (group_type == TAG_TO_READER && d->use_tag) manch_step(d, (double)e->v, dur);
    else if (group_type == READER_TO_TAG && d->use_reader) miller_step(d, (double)e->v, dur);
}

void nfc_dec_feed(nfc_dec *d, const nfc_event *ev, int64_t n, double factor) {
    /* background.py:42-52.  The reference collects maximal same-type groups and calls
     * process_transition once per group; the decoders keep their state between
     * calls, so stepping event by event inside each group is the same computation. */
    int cur = TAG_TO_READER;
    int64_t gstart = 0;
    for (int64_t i = 0; i < n; i++) {
        if (ev[i].type != cur) {
            for (int64_t k = gstart; k < i; k++) dec_one(d, &ev[k], factor, cur);
            gstart = i;
            cur = ev[i].type;
        }
    }
    for (int64_t k = gstart; k < n; k++) dec_one(d, &ev[k], factor, cur);
}

const nfc_symbol *nfc_dec_symbols(const nfc_dec *d, int64_t *count) {
    *count = d->sym.n;
    return d->sym.p;
}
const nfc_frame *nfc_dec_frames(const nfc_dec *d, int64_t *count) {
    *count = d->fr.n;
    return d->fr.p;
}
const uint8_t *nfc_dec_bits(const nfc_dec *d, int64_t *count) {
    *count = d->bits.n;
    return d->bits.p;
}

void nfc_dec_get_state(const nfc_dec *d, int *miller_s, int *manch_s, int *started, int *npending,
                       uint8_t *pending_bits, int pending_cap) {
    *miller_s = ml_get_stage(&d->ml) | (d->ml.has_started << 2) | ((d->ml.prev != 0) << 3);
    *manch_s = d->mc.prev_set | (((int)d->mc.prev + 1) << 1);
    int o = 0;
    for (int t = 0; t < 2; t++) {
        started[t] = d->pp[t].started;
        npending[t] = (int)d->pp[t].cur.n;
        for (int64_t i = 0; i < d->pp[t].cur.n && o < pending_cap; i++) pending_bits[o++] = d->pp[t].cur.p[i];
    }
}

void nfc_dec_set_state(nfc_dec *d, int miller_s, int manch_s, const int *started, const int *npending,
                       const uint8_t *pending_bits) {
    ml_set_stage(&d->ml, ST_BEGINNING);
    ml_set_stage(&d->ml, miller_s & 3);
    d->ml.has_started = (miller_s >> 2) & 1;
    d->ml.prev = (miller_s >> 3) & 1;
    d->mc.prev_set = manch_s & 1;
    d->mc.prev = ((manch_s >> 1) & 3) - 1;
    int o = 0;
    for (int t = 0; t < 2; t++) {
        d->pp[t].started = started[t];
        d->pp[t].cur.n = 0;
        for (int i = 0; i < npending[t]; i++) bytevec_push(&d->pp[t].cur, pending_bits[o++]);
    }
}

/* ------------------------------------------------------------- fsm.py tail */
int32_t nfc_fix_ending(const uint8_t *bits, int32_t n, int type, uint8_t *out, int *flag) {
    /* fsm.py:51-66 */
    int rem = n % 9;
    int sb = start_bit(type);
    *flag = 0;
    if (rem == 0) {
        memcpy(out, bits, (size_t)n);
        return n;
    } else if (rem == 8) {
        memcpy(out, bits, (size_t)n);
        out[n] = (uint8_t)sb;
        return n + 1;
    } else if (rem == 1) {
        if (bits[n - 1] != sb) *flag = 1;
        memcpy(out, bits, (size_t)(n - 1));
        return n - 1;
    }
    *flag = 2;
    memcpy(out, bits, (size_t)(n - rem));
    return n - rem;
}

int32_t nfc_check_parity(const uint8_t *bits, int32_t n, uint8_t *bytes_out) {
    /* fsm.py:28-49 */
    int nb = 0, cur_byte = 0, set_bits = 0, cur_ind = 0;
    for (int i = 0; i < n; i++) {
        int bit = bits[i];
        if (cur_ind < 8) {
            cur_byte |= (bit << cur_ind);
            cur_ind += 1;
            set_bits += bit;
        } else {
            if ((set_bits & 1) == bit) return -1;
            bytes_out[nb++] = (uint8_t)cur_byte;
            set_bits = cur_byte = cur_ind = 0;
        }
    }
    if (cur_ind == 8) bytes_out[nb++] = (uint8_t)cur_byte;
    return nb;
}

int32_t nfc_print_enc(const uint8_t *bits, int32_t n, uint8_t *bytes_out, uint8_t *flag_out) {
    /* fsm.py:114-131 */
    int nb = 0, cur_byte = 0, set_bits = 0, cur_ind = 0;
    for (int i = 0; i < n; i++) {
        int bit = bits[i];
        if (cur_ind < 8) {
            cur_byte |= (bit << cur_ind);
            cur_ind += 1;
            set_bits += bit;
        } else {
            flag_out[nb] = (uint8_t)((set_bits & 1) == bit);
            bytes_out[nb++] = (uint8_t)cur_byte;
            set_bits = cur_byte = cur_ind = 0;
        }
    }
    return nb;
}

void nfc_crc_a(const uint8_t *data, int32_t n, uint8_t out[2]) {
    /* utilities.py:26-41, cktp = CRC_14443_A */
    uint32_t wcrc = 0x6363;
    for (int32_t i = 0; i < n; i++) {
        uint32_t b = data[i];
        b = b ^ (wcrc & 0xFF);
        b = b ^ ((b << 4) & 0xFF); /* `b ^ (b << 4) & 0xFF`: & binds tighter than ^ */
        wcrc = (wcrc >> 8) ^ (b << 8) ^ (b << 3) ^ (b >> 4);
    }
    out[0] = (uint8_t)(wcrc & 0xFF);
    out[1] = (uint8_t)((wcrc >> 8) & 0xFF);
}

int nfc_check_crc(const uint8_t *data, int32_t n) {
    /* utilities.py:43-46; the reference indexes data[-2], data[-1]: n >= 2 */
    uint8_t crc[2];
    if (n < 2) return 0;
    nfc_crc_a(data, n - 2, crc);
    return crc[0] == data[n - 2] && crc[1] == data[n - 1];
}

/* ----------------------------------------------------------------- encoders */
int32_t nfc_miller_encode(const uint8_t *bits, int32_t n, int8_t *level, double *dur_us) {
    /* miller.py:200-233 */
    static const int8_t ONE_l[3] = {1, 0, 1};
    const double ONE_d[3] = {PL_HALF, PL_ZERO, PL_ONE_REM};
    static const int8_t ZERO0_l[2] = {0, 1};
    const double ZERO0_d[2] = {PL_ZERO, PL_ZERO_REM};
    static const int8_t ZERO1_l[1] = {1};
    const double ZERO1_d[1] = {PL_FULL};
    int32_t m = 0;
    level[0] = ZERO0_l[0];
    dur_us[0] = ZERO0_d[0];
    level[1] = ZERO0_l[1];
    dur_us[1] = ZERO0_d[1];
    m = 2;
    int last_bit = 0;
    for (int32_t i = 0; i <= n; i++) {
        int bit = i < n ? bits[i] : 0; /* appended 0 signifies end, :209 */
        const int8_t *cl = ONE_l;
        const double *cd = ONE_d;
        int cn = 3;
        if (bit == 0) {
            if (last_bit == 0) { cl = ZERO0_l; cd = ZERO0_d; cn = 2; }
            else { cl = ZERO1_l; cd = ZERO1_d; cn = 1; }
        }
        last_bit = bit;
        int k0 = 0;
        if (cl[0] == level[m - 1]) {
            dur_us[m - 1] = cd[0] + dur_us[m - 1]; /* cur_dur = start_dur + last_dur */
            k0 = 1;
        }
        for (int k = k0; k < cn; k++) {
            level[m] = cl[k];
            dur_us[m] = cd[k];
            m++;
        }
    }
    return m;
}

int32_t nfc_manchester_encode(const uint8_t *bits, int32_t n, int8_t *level, double *dur_us) {
    /* manchester.py:64-79 */
    int32_t m = 0;
    level[m] = 1; dur_us[m++] = PL_HALF;
    level[m] = 0; dur_us[m++] = PL_HALF;
    int last = 0;
    for (int32_t i = 0; i < n; i++) {
        int bit = bits[i];
        if (bit == last) {
            level[m - 1] = (int8_t)bit;
            dur_us[m - 1] = PL_FULL;
            last = 1 - last;
            level[m] = (int8_t)last; dur_us[m++] = PL_HALF;
        } else {
            level[m] = (int8_t)(1 - last); dur_us[m++] = PL_HALF;
            level[m] = (int8_t)last; dur_us[m++] = PL_HALF;
        }
    }
    return m;
}

/* -------------------------------------------------------------- whole chain */
static uint64_t fnv(uint64_t h, const void *p, size_t n) {
    const uint8_t *b = (const uint8_t *)p;
    for (size_t i = 0; i < n; i++) {
        h ^= b[i];
        h *= 1099511628211ULL;
    }
    return h;
}

void nfc_chain_run(const float *in, int64_t n, double samp_rate, double lo_val, double hi_val,
                   int av_window, int max_len, int decode_reader, int decode_tag, int64_t chunk,
                   nfc_chain_result *res) {
    nfc_ts *t = nfc_ts_new(samp_rate, lo_val, hi_val, av_window, max_len);
    nfc_dec *d = nfc_dec_new(decode_reader, decode_tag);
    uint64_t h = 1469598103934665603ULL;
    int64_t nev = 0;
    int64_t off = 0;
    if (chunk <= 0) chunk = 8192;
    while (off < n) {
        int64_t m = n - off < chunk ? n - off : chunk;
        int cb = 0;
        int64_t used = nfc_ts_work(t, in + off, m, &cb);
        if (cb) {
            int64_t c;
            const nfc_event *ev = nfc_ts_events(t, &c);
            for (int64_t i = 0; i < c; i++) {
                h = fnv(h, &ev[i].pos, 8);
                h = fnv(h, &ev[i].d, 4);
                h = fnv(h, &ev[i].v, 1);
                h = fnv(h, &ev[i].type, 1);
            }
            nev += c;
            nfc_dec_feed(d, ev, c, t->factor);
        }
        off += used;
        if (used == 0 && m == 0) break;
    }
    for (int64_t i = 0; i < d->sym.n; i++) {
        h = fnv(h, &d->sym.p[i].pos, 8);
        h = fnv(h, &d->sym.p[i].type, 1);
        h = fnv(h, &d->sym.p[i].val, 1);
    }
    for (int64_t i = 0; i < d->fr.n; i++) {
        h = fnv(h, &d->fr.p[i].pos, 8);
        h = fnv(h, &d->fr.p[i].type, 4);
        h = fnv(h, &d->fr.p[i].nbits, 4);
        h = fnv(h, d->bits.p + d->fr.p[i].bit_off, (size_t)d->fr.p[i].nbits);
    }
    res->n_events = nev;
    res->n_symbols = d->sym.n;
    res->n_frames = d->fr.n;
    res->n_bits = d->bits.n;
    res->digest = h;
    nfc_dec_free(d);
    nfc_ts_free(t);
}
