"""GNU Radio base classes when GNU Radio is installed, minimal stand-ins when it is not.

The reference's blocks derive from gr.sync_block / gr.hier_block2 (transition_sink.py:10,
decoder.py:15).  GNU Radio is absent from the build and test machines; the stand-ins keep the
constructor signature (name, in_sig, out_sig) so that the blocks can be driven by calling work().
"""
try:  # pragma: no cover - exercised only where GNU Radio exists
    from gnuradio import gr  # noqa: F401
    from gnuradio import blocks  # noqa: F401
    HAVE_GNURADIO = True
except Exception:  # ImportError, or a broken installation
    HAVE_GNURADIO = False
    blocks = None

    class _Block(object):
        def __init__(self, name=None, in_sig=None, out_sig=None, *a, **k):
            self._name, self._in_sig, self._out_sig = name, in_sig, out_sig

        def name(self):
            return self._name

        def connect(self, *a, **k):
            pass

    class gr(object):  # noqa: N801 - mirrors the module name
        sync_block = _Block
        hier_block2 = _Block
        top_block = _Block

        @staticmethod
        def io_signature(*a, **k):
            return None
