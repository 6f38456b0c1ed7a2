"""f3ps: Python mirror of the B200-native supervoxel-plus-merging path (libf3ps.so, include/f3ps.h)."""
from .binding import (Segmenter, merge_batch, F3psError, LogicError, build, lib, LIB_PATH, EXPORTED,
                      LAB_CIEDE00, RGB_EUCL, NORMALS_DIFF, CONVEX_NORMALS_DIFF,
                      MANUAL_LAMBDA, ADAPTIVE_LAMBDA, EQUALIZATION)
from . import synth
