"""Drop-in replacement of the reference's transition_sink block (transition_sink.py:10-125).

Same class name, constructor parameters, defaults, item type and callback contract; the per-sample
loop runs on the GPU behind the C ABI.  The callback receives [((v, dur_us), type), ...] exactly once
per work() call after the warm-up, even when the list is empty (transition_sink.py:101).
"""
import numpy

from . import _cabi
from .grcompat import gr


class transition_sink(gr.sync_block):
    "Transition sink"

    def __init__(self, samp_rate, callback, lo_val=0.1, hi_val=1.1, av_window=2000, max_len=50, device=0):
        gr.sync_block.__init__(
            self,
            name="transition_sink",
            in_sig=[numpy.float32],
            out_sig=None,
        )
        self._callback = callback
        self._factor = 1e6 / samp_rate
        self._stream = _cabi.Stream(samp_rate, lo_val, hi_val, av_window, max_len, reader=False, tag=False,
                                    input_kind=_cabi.IN_ENVELOPE_F32,
                                    outputs=_cabi.OUT_EVENTS | _cabi.OUT_DROPPED_EVENTS, device=device)

    def work(self, input_items, output_items):
        consumed, called_back = self._stream.push(input_items[0])
        if called_back:
            ev = self._stream.drain_events()
            factor = self._factor
            self._callback([((int(v), int(d) * factor), int(t)) for v, d, t in zip(ev["v"], ev["d"], ev["type"])])
        return consumed

    def stream(self):
        return self._stream
