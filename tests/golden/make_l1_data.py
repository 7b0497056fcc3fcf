"""Writes tests/golden/l1_data.npz from the reference build (oracle/_ref/libgmr1_ref.so = the reference's
src/l1/conv.c, crc.c, punct.c compiled verbatim): every convolutional-code table, the CRC parameters, the 51
puncturing masks and the puncturing arrays gmr1_puncturer_generate makes for tests/l1_data.py:GENERATE.
Run in the build container (needs oracle/_ref); the output is committed."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import l1_data

ref = os.path.join(l1_data.ROOT, "oracle", "_ref", "libgmr1_ref.so")
d = l1_data.read_all(ref)
np.savez_compressed(os.path.join(HERE, "l1_data.npz"), **d)
print(len(d), "arrays")
