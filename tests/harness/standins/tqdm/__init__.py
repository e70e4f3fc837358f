"""Harness stand-in for ``tqdm``: run_nerf.py hard-codes ``N_iters = 50000 + 1`` (:536) and offers no option to
stop earlier.  ``trange`` here yields at most HN_HARNESS_ITERS iterations, which is how the harness bounds an
UNMODIFIED run_nerf.py; ``tqdm`` passes its iterable through.  TEST HARNESS ONLY."""
import os
import sys


def trange(*args, **_kw):
    limit = int(os.environ.get("HN_HARNESS_ITERS", "50"))
    for n, i in enumerate(range(*args)):
        if n >= limit:
            return
        yield i


class tqdm:
    def __init__(self, iterable=None, *a, **k):
        self.iterable = iterable

    def __iter__(self):
        return iter(self.iterable)

    @staticmethod
    def write(msg, *a, **k):
        sys.stdout.write(str(msg) + "\n")
