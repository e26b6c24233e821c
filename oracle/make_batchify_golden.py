"""TEST INFRASTRUCTURE ONLY -- golden partitions from the UNMODIFIED reference batching policies.

    python oracle/make_batchify_golden.py      (dev container only: needs /root/reference)

Loads ``liteasr/utils/batchify.py`` of the reference through the shims of ``oracle/ref_shims.py`` (its three imports --
``liteasr.config``, ``liteasr.dataclass.audio_data`` (soundfile), ``liteasr.utils.progress_bar`` -- are stubbed; not one line
of the policies themselves is replaced), runs ``SeqBatch`` / ``FrameBatch`` on seeded length lists in the order
``dataset/asr_dataset.py:107-110`` produces, and writes ``tests/golden/batchify.json``.
"""
from __future__ import annotations

import json
import os
import random
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402


def load_reference_policies():
    ref_shims.install()

    class _Bar:
        def __init__(self, *a, **k):
            pass

        def update(self, *a, **k):
            pass

    ref_shims._stub("liteasr.dataclass", __path__=[])
    ref_shims._stub("liteasr.dataclass.audio_data", Audio=object)
    ref_shims._stub("liteasr.utils.progress_bar", ProgressBar=_Bar)
    import importlib
    return importlib.import_module("liteasr.utils.batchify")


class Sample:
    def __init__(self, xlen, ylen):
        self.xlen, self.ylen = xlen, ylen


def cases():
    rng = random.Random(42)
    out = []
    for n, cfg in [
        (57, dict(batch_count="seq", batch_size=8, min_batch_size=1, max_len_in=400, max_len_out=20)),
        (200, dict(batch_count="seq", batch_size=32, min_batch_size=4, max_len_in=512, max_len_out=150)),
        (31, dict(batch_count="seq", batch_size=4, min_batch_size=0, max_len_in=100, max_len_out=10)),      # quirk Q-a
        (1, dict(batch_count="seq", batch_size=16, min_batch_size=1, max_len_in=800, max_len_out=150)),
        (120, dict(batch_count="frame", max_frame_in=6000, max_frame_out=0, max_frame_inout=0)),
        (120, dict(batch_count="frame", max_frame_in=0, max_frame_out=300, max_frame_inout=0)),
        (150, dict(batch_count="frame", max_frame_in=9000, max_frame_out=400, max_frame_inout=9200)),
        (40, dict(batch_count="frame", max_frame_in=900, max_frame_out=0, max_frame_inout=0)),             # quirk Q-b
        (0, dict(batch_count="frame", max_frame_in=900, max_frame_out=0, max_frame_inout=0)),
    ]:
        xl = [rng.randint(50, 1200) for _ in range(n)]
        if n > 10:
            xl[3] = xl[7] = xl[8]  # ties: the sort must be stable
        yl = [max(1, min(x // 8, rng.randint(2, 60))) for x in xl]
        out.append((xl, yl, cfg))
    return out


def main():
    ref = load_reference_policies()
    golden = []
    for xl, yl, cfg in cases():
        full = dict(batch_size=None, min_batch_size=None, max_len_in=None, max_len_out=None, max_frame_in=None,
                    max_frame_out=None, max_frame_inout=None)
        full.update(cfg)
        c = types.SimpleNamespace(**full)
        samples = [Sample(x, y) for x, y in zip(xl, yl)]
        pol = (ref.SeqBatch if c.batch_count == "seq" else ref.FrameBatch)(c)
        if samples:  # dataset/asr_dataset.py:107-110
            indices, _ = zip(*sorted(enumerate(samples), key=lambda d: d[1].xlen, reverse=True))
        else:
            indices = ()
        pol.batchify(indices, samples)
        golden.append(dict(xlens=xl, ylens=yl, cfg=full, order=list(indices), batches=[list(b) for b in pol.data]))
    path = os.path.join(ROOT, "tests", "golden", "batchify.json")
    with open(path, "w") as f:
        json.dump(golden, f)
    print(path, [len(g["batches"]) for g in golden], "empty batches:", [sum(1 for b in g["batches"] if not b) for g in golden])


if __name__ == "__main__":
    main()
