"""usrp_nfc_b200 -- the sample-rate decode path of giech/usrp_nfc on NVIDIA B200 (sm_100a).

Public surface (mirrors the reference's block names):
    transition_sink.transition_sink   drop-in for code/transition_sink.py
    decoder.decoder                   drop-in for code/decoder.py (fused GPU path up to fsm.process_bits)
    _cabi.Stream                      the C ABI of include/usrp_nfc_b200.h, one handle per stream
    sharding                          time-segment / batch sharding across GPUs (torch.distributed)
    synth                             synthetic ISO 14443A traffic for tests and benchmarks
Importing the package does not load the CUDA library; the first use does and fails loudly if it is missing.
"""
__all__ = ["transition_sink", "decoder", "synth", "_cabi"]
