"""TEST / INTEGRATION EXAMPLE, not product code: a Python mirror of the reference's trial dispatch (src/evaluate.rs:24-201)
over the fused GPU filter call.  In a real integration this control flow stays in the reference's Rust;
it lives under tests/ so that the flow tests and bench.py's -o2 leg can drive the C ABI the way the reference would.

The reference fans out `filters.par_iter()` -> `image.filter_image(filter, optimize_alpha)` -> `deflater.deflate(...)`
on rayon workers (evaluate.rs:141-154).  Here all strategies of a candidate go to the GPU in ONE call
(`oxi_filter_image_multi`: the image is uploaded once and comes back as K filtered streams) and only the deflate fan-out
stays on host threads -- exactly the split the north star asks for (deflate and Brute stay on the host).

The reference deflates with libdeflater; this mirror uses zlib as the stand-in so that the flow is runnable here.  The
ranking rule (`Candidate::cmp_key`, evaluate.rs:38-46) and the "skip if larger than the best known size" rule
(:165-169) are the reference's.
"""
from __future__ import annotations

import concurrent.futures as cf
import zlib
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from oxipng_b200.png import FilterStrategy, PngImage, filter_image_multi

_VARIANT = {"Basic": 0, "MinSum": 1, "Entropy": 2, "Bigrams": 3, "BigEnt": 4, "Brute": 5, "Predefined": 6}


def _strategy_key(s: FilterStrategy):  # derive(Ord) on FilterStrategy, strategies.rs:8-29
    return (_VARIANT[s.kind], int(s.filter) if s.filter is not None else 0, tuple(s.lines) if s.lines else ())


def key_chunks_size(png: PngImage) -> int:  # png/mod.rs:305-320
    ct = png.ihdr.color_type
    if ct.code == 3:
        pal = ct.palette_array()
        plte = 12 + len(pal) * 3
        nz = np.flatnonzero(pal[:, 3] != 255) if len(pal) else np.array([], dtype=np.int64)
        return plte if nz.size == 0 else plte + 12 + int(nz[-1]) + 1
    if ct.code == 0 and ct.transparent is not None:
        return 12 + 2
    if ct.code == 2 and ct.transparent is not None:
        return 12 + 6
    return 0


@dataclass
class Candidate:  # evaluate.rs:24-47
    image: PngImage
    idat_data: Optional[bytes]
    estimated_output_size: int
    filter: FilterStrategy
    filter_used: FilterStrategy
    nth: int

    def cmp_key(self):
        return (self.estimated_output_size, self.image.data.size, _strategy_key(self.filter), (1 << 64) - 1 - self.nth)


class Evaluator:  # evaluate.rs:50-201
    def __init__(self, filters: Sequence[FilterStrategy], zlib_level: int = 6, optimize_alpha: bool = False,
                 final_round: bool = False, threads: int = 8):
        self.filters = list(dict.fromkeys(filters))  # IndexSet
        self.zlib_level = zlib_level
        self.optimize_alpha = optimize_alpha
        self.final_round = final_round
        self.nth = 0
        self.best_candidate_size: Optional[int] = None
        self._candidates: List[Candidate] = []
        self._pool = cf.ThreadPoolExecutor(max_workers=threads)

    def set_best_size(self, size: int) -> None:  # evaluate.rs:115-117
        self.best_candidate_size = size if self.best_candidate_size is None else min(self.best_candidate_size, size)

    def try_image(self, image: PngImage, description: str = "") -> None:  # evaluate.rs:120-201
        nth = self.nth
        self.nth += 1
        # one fused GPU call for every strategy of this candidate (evaluate.rs:143-153 as a single launch)
        if hasattr(image, "filter_image_multi"):  # device-resident candidate (oxi_image_filter): only streams come back
            outs, used = image.filter_image_multi(self.filters, self.optimize_alpha)
        else:
            outs, used = filter_image_multi(image, self.filters, self.optimize_alpha)
        kcs = key_chunks_size(image)

        def deflate(j):
            return zlib.compress(outs[j].tobytes(), self.zlib_level)

        for j, idat in enumerate(self._pool.map(deflate, range(len(self.filters)))):
            f = self.filters[j]
            est = len(idat) + kcs  # estimated_output_size, png/mod.rs:323-326
            if self.best_candidate_size is not None and est > self.best_candidate_size:
                continue  # evaluate.rs:165-169
            heuristic = f.kind not in ("Basic", "Predefined") and used[j].size > 0
            fu = FilterStrategy.Predefined(used[j].tolist()) if heuristic else f
            self._candidates.append(Candidate(image, idat if self.final_round else None, est, f, fu, nth))
            self.set_best_size(est)

    def get_best_candidate(self) -> Optional[Candidate]:  # evaluate.rs:96-112
        self._pool.shutdown()
        return min(self._candidates, key=Candidate.cmp_key) if self._candidates else None
