"""Development aid: run the stage kernel of the `trace` build variant once and dump the per-warp clock stamps
(events 0..7 of 4 planes of 296 CTAs) to gpurun_out/pipe_trace.npz.  See LK_PIPE_TRACE in lk_pipe.cuh."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

here = os.path.dirname(os.path.abspath(__file__))
lib = os.path.join(os.path.dirname(here), "loki_b200", "libloki_b200_trace.so")
os.environ["LOKI_B200_LIB"] = lib
sys.path.insert(0, os.path.dirname(here))
sys.argv = [sys.argv[0], "--reps", "1"]
import microbench_rhs  # noqa: E402
microbench_rhs.main()
import loki_b200  # noqa: E402
L = loki_b200.load()
st = np.zeros(296 * 8 * 4 * 8, dtype=np.int64)
sm = np.zeros(296, dtype=np.int32)
assert L.lk_debug_pipe_trace(st.ctypes.data_as(C.c_void_p), sm.ctypes.data_as(C.c_void_p)) == 0
np.savez(os.path.join(os.path.dirname(here), "gpurun_out", "pipe_trace.npz"), stamps=st.reshape(296, 8, 4, 8), smid=sm)
print("trace saved")
