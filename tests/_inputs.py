"""Test inputs that are not one of the four synthetic families: real (non-synthetic) source text found in the image."""
import glob
import os

import numpy as np


def real_source_text(nbytes):
    """Python source files of the interpreter's standard library and site-packages, concatenated (sorted order, so the
    same box always yields the same bytes) — LCP-heavy, sigma ~ 100-200: the realistic counterpart of BASELINE config 3."""
    chunks, have = [], 0
    for root in (os.path.dirname(os.__file__), os.path.dirname(os.path.dirname(np.__file__))):
        for path in sorted(glob.glob(os.path.join(root, "**", "*.py"), recursive=True)):
            try:
                b = open(path, "rb").read()
            except OSError:
                continue
            chunks.append(np.frombuffer(b, np.uint8))
            have += len(b)
            if have >= nbytes:
                break
        if have >= nbytes:
            break
    data = np.concatenate(chunks) if chunks else np.zeros(0, np.uint8)
    return data[:nbytes].copy()
