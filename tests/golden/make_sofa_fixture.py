"""Generates tests/golden/cipic003_hrir.npz from the reference's bundled SOFA file
(/root/reference/data/hrtf/subject_003.sofa, CIPIC subject 003, SimpleFreeFieldHRIR, M=1250 x R=2 x N=200, 44.1 kHz).

Run in the dev container only (the reference tree does not exist on the GPU box); the .npz is committed.
No HDF5 library is needed: the file holds two zlib streams (SourcePosition [1250][3] f64 and Data.IR [1250][2][200]
f64), each HDF5 byte-shuffled.  They are located by scanning for zlib headers and checking the decoded size
(SURVEY.md §8c).  All IR values are exactly representable in f32 (asserted).
Also converts data/hrtf/processed_hrir.wav (4 ch x 200 frames, int16) to f32 [4][200] / 32768.
"""
import os
import struct
import zlib

import numpy as np

REF = "/root/reference/data/hrtf"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cipic003_hrir.npz")


def find_stream(buf: bytes, decoded_size: int) -> bytes:
    i = 0
    while True:
        i = buf.find(b"\x78", i)
        if i < 0:
            raise RuntimeError("zlib stream with decoded size %d not found" % decoded_size)
        if buf[i + 1] in (0x01, 0x5E, 0x9C, 0xDA):
            try:
                d = zlib.decompressobj()
                out = d.decompress(buf[i:])
                if len(out) == decoded_size:
                    return out
            except zlib.error:
                pass
        i += 1


def unshuffle_f64(raw: bytes, shape):
    a = np.frombuffer(raw, np.uint8).reshape(8, -1).T.copy()
    return a.view("<f8").reshape(shape)


def main():
    buf = open(os.path.join(REF, "subject_003.sofa"), "rb").read()
    pos = unshuffle_f64(find_stream(buf, 1250 * 3 * 8), (1250, 3))
    ir = unshuffle_f64(find_stream(buf, 1250 * 2 * 200 * 8), (1250, 2, 200))
    ir32 = ir.astype(np.float32)
    assert np.array_equal(ir32.astype(np.float64), ir), "IR not exactly f32-representable"
    assert abs(pos[308, 0] - 30.0) < 1e-9 and abs(pos[308, 1]) < 1e-9, pos[308]
    assert abs(pos[908, 0] - 330.0) < 1e-9 and abs(pos[908, 1]) < 1e-9, pos[908]
    assert abs(pos[608, 0]) < 1e-9 and abs(pos[608, 1]) < 1e-9, pos[608]

    wav = open(os.path.join(REF, "processed_hrir.wav"), "rb").read()
    di = wav.find(b"data")
    n = struct.unpack("<I", wav[di + 4:di + 8])[0]
    pcm = np.frombuffer(wav[di + 8:di + 8 + n], "<i2").reshape(-1, 4).T
    processed = (pcm.astype(np.float32) / 32768.0).copy()

    np.savez_compressed(OUT, ir=ir32, pos=pos.astype(np.float32), processed_hrir=processed, fs=np.float32(44100.0))
    print("wrote", OUT, os.path.getsize(OUT), "bytes; ir", ir32.shape, "max", float(np.abs(ir32).max()),
          "processed", processed.shape)


if __name__ == "__main__":
    main()
