// tables.h -- host-side line-code transition tables (see tables.cpp).
#pragma once
#include <vector>

#include "common.cuh"

namespace nfc {

struct HostTables {
    int max_len = 0;
    int n_dclass_miller = 0, n_dclass_manch = 0;
    std::vector<uint8_t> dclass_miller, dclass_manch;  // [d], d = 0..max_len
    std::vector<TabEntry> miller;                       // [dclass][v + 1][state], 16 states
    std::vector<TabEntry> manch;                        // [dclass][v + 1][state], 8 states
    // [dclass][v + 1]: 0xFF, or the machine state (decoder state | _started << 4 resp. << 3) every state is
    // sent to by such an event -- an event whose duration is out of range resets the decoder and closes the frame
    std::vector<uint8_t> reset_miller, reset_manch;
};

// factor = 1e6 / samp_rate (transition_sink.py:21).  Returns false if the decoders distinguish more
// than MAX_DCLASS duration classes at this rate (does not happen for max_len <= 65535).
bool build_tables(int max_len, double factor, HostTables &t);

}  // namespace nfc
