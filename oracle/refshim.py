"""Import the UNMODIFIED reference modules from /root/reference/code -- TEST INFRASTRUCTURE.

Used by oracle/gen_golden.py (in the build container, where /root/reference exists) to
produce tests/golden/.  Nothing here is copied from the reference: the reference files
are loaded from where they lie.

  * transition_sink.py, miller.py, manchester.py, utilities.py, cipher.py, lfsr.py load
    verbatim under Python 3 once a stub `gnuradio` package is importable (and
    builtins.xrange = range).
  * packets.py, fsm.py, command.py contain Python 2 print statements and a handful of
    integer `/`; they are transformed IN MEMORY at import time (print -> function,
    xrange -> range, `)/2` and `)/9` -> `//`), which changes no arithmetic.
"""
import importlib.abc
import importlib.util
import io
import os
import re
import sys
import types

REF_CODE = os.environ.get("USRP_NFC_REFERENCE", "/root/reference/code")
_VERBATIM = {"transition_sink", "miller", "manchester", "utilities", "cipher", "lfsr"}
_TRANSFORMED = {"packets", "fsm", "command"}

log = io.StringIO()  # everything the reference prints goes here


def available():
    return os.path.isfile(os.path.join(REF_CODE, "transition_sink.py"))


def _p2print(*args, **kw):
    end = kw.get("end", "\n")
    log.write(" ".join(str(a) for a in args) + end)


def _transform(src):
    out = []
    for line in src.splitlines():
        m = re.match(r"^(\s*)print\b\s?(.*)$", line)
        if m and not line.lstrip().startswith("#"):
            indent, rest = m.group(1), m.group(2).rstrip()
            if rest.endswith(","):
                line = "%s_p2print(%s end=' ')" % (indent, rest)
            elif rest == "":
                line = "%s_p2print()" % indent
            else:
                line = "%s_p2print(%s)" % (indent, rest)
        line = line.replace("xrange(", "range(")
        line = re.sub(r"\)/(2|9)\b", r")//\1", line)
        out.append(line)
    return "\n".join(out) + "\n"


class _Sync(object):
    """Stand-in for gr.sync_block / hier_block2 / top_block: accepts any constructor arguments."""

    def __init__(self, *a, **k):
        pass

    def connect(self, *a, **k):
        pass


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path=None, target=None):
        if name in _VERBATIM or name in _TRANSFORMED:
            return importlib.util.spec_from_loader(name, self)
        return None

    def create_module(self, spec):
        return None

    def exec_module(self, module):
        name = module.__name__
        path = os.path.join(REF_CODE, name + ".py")
        with open(path) as f:
            src = f.read()
        if name in _TRANSFORMED:
            src = _transform(src)
            module.__dict__["_p2print"] = _p2print
        module.__file__ = path
        exec(compile(src, path, "exec"), module.__dict__)


_installed = False


def install():
    """Make `import transition_sink, miller, manchester, packets, fsm, ...` resolve to the reference."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference checkout not found at %s" % REF_CODE)
    import builtins
    builtins.xrange = range  # cipher.py / lfsr.py / utilities.py call xrange at run time
    gnuradio = types.ModuleType("gnuradio")
    gr = types.ModuleType("gnuradio.gr")
    gr.sync_block = gr.hier_block2 = gr.top_block = _Sync
    gr.io_signature = lambda *a, **k: None
    gnuradio.gr = gr
    for sub in ("blocks", "analog", "uhd", "eng_option"):
        mod = types.ModuleType("gnuradio." + sub)
        setattr(gnuradio, sub, mod)
        sys.modules["gnuradio." + sub] = mod
    sys.modules["gnuradio"] = gnuradio
    sys.modules["gnuradio.gr"] = gr
    sys.meta_path.insert(0, _Finder())
    _installed = True


class ReferenceChain(object):
    """transition_sink -> background grouping -> manchester/miller -> packets -> fsm, all reference code.

    background.run (background.py:37-52) is a busy-wait daemon thread; its body is
    restated synchronously here (it is 10 lines of list slicing, no arithmetic).
    """

    def __init__(self, samp_rate, reader=True, tag=True, with_fsm=True, **ts_kwargs):
        install()
        import manchester
        import miller
        import packets
        import transition_sink

        self.events, self.symbols, self.frames = [], [], []
        chain = self

        class _Cpp(packets.CombinedPacketProcessor if with_fsm else object):
            def __init__(cpp):
                if with_fsm:
                    packets.CombinedPacketProcessor.__init__(cpp)
                else:
                    cpp._packet_processors = [packets.PacketProcessor(i) for i in range(packets.PacketType.NUM_TYPES)]

            def append_bit(cpp, bit, packet_type):
                chain.symbols.append((chain._cur_pos, packet_type, int(bit)))
                ret = cpp._packet_processors[packet_type].append_bit(bit)
                if ret:
                    chain.frames.append((chain._cur_pos, packet_type, [int(b) for b in ret]))
                    if with_fsm:
                        cpp._fsm.process_bits(ret, packet_type)

        self._cpp = _Cpp()
        self._reader = miller.miller_decoder(self._cpp) if reader else None
        self._tag = manchester.manchester_decoder(self._cpp) if tag else None
        self._ptype = packets.PacketType
        self._cur_pos = -1
        self.sink = transition_sink.transition_sink(samp_rate, self._on_events, **ts_kwargs)
        self._pos = 0
        self._batch_base = 0

    def _process(self, a, t):
        # background.py:30-35
        if t == self._ptype.TAG_TO_READER and self._tag:
            for tr, pos in a:
                self._cur_pos = pos
                self._tag.process_transition([tr])
        elif t == self._ptype.READER_TO_TAG and self._reader:
            for tr, pos in a:
                self._cur_pos = pos
                self._reader.process_transition([tr])

    def _on_events(self, transitions):
        # positions are recovered by the caller (see run); background.py:42-52 grouping:
        self._last_batch = transitions

    def run(self, samples, chunk=8192):
        """Feed float32 samples through work() one sample at a time to recover per-event
        positions (the reference's event stream is independent of chunking), grouped into
        `chunk`-sized batches for the background grouping."""
        import numpy as np
        x = np.asarray(samples, dtype=np.float32)
        off = 0
        n = x.size
        while off < n:
            m = min(chunk, n - off)
            batch = []
            i = 0
            while i < m:
                self._last_batch = None
                used = self.sink.work([x[off + i: off + i + 1]], None)
                if self._last_batch:
                    for tr in self._last_batch:
                        batch.append((tr, off + i))
                i += used if used else 1
            # one background batch per chunk
            a, cur = [], self._ptype.TAG_TO_READER
            for (val, t), pos in batch:
                self.events.append((pos, int(val[0]), float(val[1]), int(t)))
                if t == cur:
                    a.append((val, pos))
                else:
                    self._process(a, cur)
                    a, cur = [(val, pos)], t
            if a:
                self._process(a, cur)
            off += m
        return self

    def run_chunked(self, samples, chunks):
        """Feed with explicit work() sizes (no per-event positions): returns list of callback batches."""
        import numpy as np
        x = np.asarray(samples, dtype=np.float32)
        off, batches, ci = 0, [], 0
        while off < x.size:
            m = chunks[ci % len(chunks)]
            ci += 1
            self._last_batch = None
            used = self.sink.work([x[off: off + m]], None)
            if self._last_batch is not None:
                batches.append([(int(v[0]), float(v[1]), int(t)) for v, t in self._last_batch])
            off += used
        return batches
