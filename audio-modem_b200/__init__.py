"""audio_modem_b200 — B200-native OFDM receive path of playok/audio-modem.

Package contents: `csrc/` (hand-written sm_100a kernels and the C-ABI library), `modem` (host-side mirror of
the reference's modem.js / StreamingReceiver interface), `shard` (stream sharding and counter gather across
GPUs).  Import as `audio_modem_b200` (see the shim at the repository root).
"""
from . import _build, _lib, shard  # noqa: F401
from ._lib import AmbError, FrameResult  # noqa: F401
from .modem import (  # noqa: F401
    FRAME_DATA, FRAME_META, Modem, StreamingReceiver, analyzeLoopback, buildDataChunkFrame, buildTransmitSignal, generateTestSignal, buildMetadataFrame, crc32, decodeChunkFrame, decodeReceivedSignal, demodulateOFDM,
    detectPreamble, detectPreambleCrossCorr, estimateChannel, estimateFrameSamples, generateChannelEstSymbol,
    generatePreambleSymbol1, modulateOFDM, preprocessSignal, result_to_object, setDevice, setOFDMConfig,
)


def build(force=False, verbose=False):
    """Compile the CUDA library in-tree (nvcc, sm_100a)."""
    return _build.build(force=force, verbose=verbose)
