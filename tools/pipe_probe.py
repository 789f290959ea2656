"""prints the pipe-overlap probe table (ipclb200_pipe_mix, modes 0-8)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pailliercryptolib_b200 import capi  # noqa: E402

capi.init(0)
names = {0: "IMAD.WIDE on all 32 warps/SM", 1: "DFMA on all 32 warps/SM",
         2: "16 warps IMAD.WIDE + 16 warps DFMA", 3: "the 16 IMAD.WIDE warps alone",
         4: "the 16 DFMA warps alone", 5: "16 warps IMAD.WIDE + 16 warps add-with-carry chains",
         6: "the 16 add-with-carry warps alone", 7: "16 warps IMAD.WIDE + 16 warps LOP3/SHF",
         8: "the 16 LOP3/SHF warps alone"}
for mode in range(9):
    print("mode %d  %8.3f ms   %s" % (mode, capi.pipe_mix(mode), names[mode]), flush=True)
