// tables.cpp -- host-side construction of the line-code transition tables.
//
// The Manchester (tag) and modified-Miller (reader) decoders of the reference
// (manchester.py:13-61, miller.py:13-197) look at an event's duration only through comparisons
// with constants, and the slicer emits only d in 1..max_len samples.  So for a given sample rate
// each decoder is a finite automaton over (state, v, d).  The rules below restate one decoder
// step; build_tables() evaluates them for every (state, v, d), then merges durations that
// behave identically into duration classes.  The device side (linecode.cu) is integer-only.
#include "tables.h"

#include <cmath>
#include <cstring>
#include <map>
#include <vector>

namespace nfc {

namespace {
// utilities.py:7-23
enum { TOO_SHORT = 2, TOO_LONG = 3, ENCODING = 4, INTERNAL = 5, WRONG_DUR = 6 };
const double FULL = 9.44, ZERO = 3.00;
const double HALF = FULL / 2, ZERO_REM = FULL - ZERO, ONE_REM = HALF - ZERO, ONE_HALF = FULL + HALF;

struct Step {
    int next;
    int nout;
    int out[2];
};

// miller state: bits 0-1 stage (0 BEGINNING, 1 ZERO_STAGE_0, 2 ONE_STAGE_0, 3 ONE_STAGE_1; miller.py:14-17),
// bit 2 _has_started, bit 3 _prev
Step miller_step(int state, int cur, double dur) {
    const double thres = 1.5, lo = ZERO - thres, hi = 2 * FULL;  // miller.py:25-27
    int stage = state & 3, started = (state >> 2) & 1, prev = (state >> 3) & 1;
    Step s = {0, 0, {0, 0}};
    auto close = [&](double d, double av) { return std::fabs(d - av) <= thres; };  // miller.py:62-63
    auto pack = [&]() { return stage | (started << 2) | (prev << 3); };

    if (cur == 0 && std::fabs(dur - ZERO) < ZERO / 2) dur = ZERO;  // miller.py:157-158
    const int cur_type = (stage == 2 || stage == 3) ? 1 : 0;       // miller.py:43-59
    int err = 0;
    if ((dur < lo || dur > hi) && (stage == 1 || stage == 3)) {  // miller.py:165-167
        s.out[s.nout++] = cur_type;
        err = TOO_LONG;
    } else if (dur < lo) {
        err = TOO_SHORT;
    } else if (dur > hi) {
        err = TOO_LONG;
    }
    if (err) {  // miller.py:173-176: _reset(); the reference leaves _prev untouched here
        s.out[s.nout++] = err;
        stage = 0;
        started = 0;
        // _prev is dead after a reset: it is read only in handle_beginning with _has_started set, which is
        // reached only through ZERO_STAGE_0, and every exit of that stage writes _prev or resets again.
        // Canonical 0 makes out-of-range events send every state to one state (frame-boundary search).
        prev = 0;
        s.next = pack();
        return s;
    }
    int rets[2], nr = 0;
    if (stage == 0) {  // handle_beginning, miller.py:73-96
        if (cur == 0) {
            if (close(dur, ZERO)) { stage = 1; started = 1; }
            else rets[nr++] = TOO_LONG;
        } else if (started) {
            const int bit = prev == 0 ? ENCODING : 0;
            if (close(dur, HALF)) stage = 2;
            else if (close(dur, FULL)) rets[nr++] = bit;
            else if (close(dur, ONE_HALF)) { rets[nr++] = bit; stage = 2; }
            else rets[nr++] = WRONG_DUR;
        }
    } else if (stage == 1) {  // handle_zs0, miller.py:98-112
        if (cur == 0) rets[nr++] = ENCODING;
        else if (close(dur, ZERO_REM)) { stage = 0; rets[nr++] = 0; }
        else if (close(dur, ZERO_REM + HALF)) { stage = 2; rets[nr++] = 0; }
        else rets[nr++] = WRONG_DUR;
    } else if (stage == 2) {  // handle_os0, miller.py:114-122
        if (cur != 0) rets[nr++] = ENCODING;
        else if (!close(dur, ZERO)) rets[nr++] = WRONG_DUR;
        else stage = 3;
    } else {  // handle_os1, miller.py:124-148
        if (cur != 1) {
            rets[nr++] = ENCODING;
        } else if (close(dur, ONE_REM)) {
            rets[nr++] = 1;
            stage = 0;
        } else {
            rets[nr++] = 1;
            stage = 0;
            dur -= ONE_REM;
            if (close(dur, FULL)) rets[nr++] = 0;
            else if (close(dur, HALF)) stage = 2;
            else if (close(dur, ONE_HALF)) { rets[nr++] = 0; stage = 2; }
            else rets[nr++] = WRONG_DUR;
        }
    }
    for (int i = 0; i < nr; i++) {  // miller.py:191-197
        s.out[s.nout++] = rets[i];
        if (rets[i] > 1) { stage = 0; started = 0; prev = 0; }
        else prev = rets[i];
    }
    s.next = pack();
    return s;
}

// manchester state: bit 0 _prev_set, bits 1-2 (_prev + 1) with _prev in {-1,0,1,2}
Step manch_step(int state, int cur, double dur) {
    const double lo = HALF - 1, mid = HALF + 1, hi = 2 * HALF + 1;  // manchester.py:17-20
    int prev_set = state & 1, prev = ((state >> 1) & 3) - 1;
    Step s = {0, 0, {0, 0}};
    // _prev is read only while _prev_set: canonical 0 otherwise
    auto pack = [&]() { return prev_set | (((prev_set ? prev : 0) + 1) << 1); };
    int err = 0;
    if (dur < lo) err = TOO_SHORT;
    else if (dur > hi) err = TOO_LONG;
    if (err) {  // manchester.py:40-43
        prev_set = 0;
        prev = 0;
        s.out[s.nout++] = err;
        s.next = pack();
        return s;
    }
    const bool dual = dur > mid;
    if (prev_set) {  // manchester.py:48-54
        if (prev == cur || (prev != 0 && prev != 1)) {
            s.out[s.nout++] = INTERNAL;
            s.next = pack();
            return s;
        }
        s.out[s.nout++] = prev;
        prev_set = dual ? 1 : 0;
    } else {  // manchester.py:55-59
        if (dual) {
            s.out[s.nout++] = ENCODING;
            s.next = pack();
            return s;
        }
        prev_set = 1;
    }
    prev = cur;
    s.next = pack();
    return s;
}

TabEntry entry_of(const Step &s) {
    return (TabEntry)((s.next & 15) | (s.nout << 4) | ((s.out[0] & 7) << 6) | ((s.out[1] & 7) << 9));
}

template <class StepFn>
bool build_one(StepFn step, int nstates, int max_len, double factor, std::vector<uint8_t> &dclass,
               std::vector<TabEntry> &table, int &nclass, std::vector<uint8_t> &reset) {
    std::map<std::vector<TabEntry>, int> seen;
    std::vector<std::vector<TabEntry>> cols;
    dclass.assign((size_t)max_len + 1, 0);
    for (int d = 0; d <= max_len; d++) {
        const double dur = (double)d * factor;  // transition_sink.py:89,97: d*factor
        std::vector<TabEntry> col((size_t)4 * nstates);
        for (int v = -1; v <= 2; v++)
            for (int st = 0; st < nstates; st++) col[(size_t)(v + 1) * nstates + st] = entry_of(step(st, v, dur));
        auto it = seen.find(col);
        int id;
        if (it == seen.end()) {
            id = (int)cols.size();
            seen.emplace(col, id);
            cols.push_back(col);
        } else {
            id = it->second;
        }
        if (id >= MAX_DCLASS) return false;
        dclass[(size_t)d] = (uint8_t)id;
    }
    nclass = (int)cols.size();
    table.assign((size_t)nclass * 4 * nstates, 0);
    for (int c = 0; c < nclass; c++) std::memcpy(&table[(size_t)c * 4 * nstates], cols[c].data(), sizeof(TabEntry) * 4 * nstates);
    // universal resets: every state goes to the same state and the last symbol emitted is an error code
    reset.assign((size_t)nclass * 4, 0xFF);
    for (int c = 0; c < nclass; c++)
        for (int v = 0; v < 4; v++) {
            const TabEntry *row = &table[((size_t)c * 4 + v) * nstates];
            bool uni = true;
            for (int st = 0; st < nstates && uni; st++) {
                const int n = tab_nout(row[st]);
                const int last = n == 0 ? 0 : (n == 1 ? tab_out0(row[st]) : tab_out1(row[st]));
                uni = tab_next(row[st]) == tab_next(row[0]) && n > 0 && last >= 2;
            }
            if (uni) reset[(size_t)c * 4 + v] = (uint8_t)tab_next(row[0]);  // _started = 0 after an error symbol
        }
    return true;
}
}  // namespace

bool build_tables(int max_len, double factor, HostTables &t) {
    t.max_len = max_len;
    if (!build_one(miller_step, MILLER_STATES, max_len, factor, t.dclass_miller, t.miller, t.n_dclass_miller, t.reset_miller)) return false;
    if (!build_one(manch_step, MANCH_STATES, max_len, factor, t.dclass_manch, t.manch, t.n_dclass_manch, t.reset_manch)) return false;
    return true;
}

}  // namespace nfc
