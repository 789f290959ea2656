"""burst vs sustained IMAD.WIDE rate, with the SM clock sampled meanwhile"""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pailliercryptolib_b200 import capi  # noqa: E402

capi.init(0)
import pynvml as nv  # noqa: E402

nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(0)
burst, mhz = capi.int_peak()
print("burst      %.3f T MAC32/s (nominal clock %.0f MHz)" % (burst / 1e12, mhz), flush=True)
os.environ["IPCLB200_PEAK_PATTERN"] = "row"
print("row pattern (16 distinct multiplicand registers per chain) %.3f T MAC32/s = %.3f of burst"
      % (capi.int_peak_sustained(0.3) / 1e12, capi.int_peak_sustained(0.3) / burst), flush=True)
os.environ.pop("IPCLB200_PEAK_PATTERN")
for sec in (0.2, 1.0):
    clk, pw, stop = [], [], [False]

    def sample():
        while not stop[0]:
            clk.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
            pw.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
            time.sleep(0.02)

    t = threading.Thread(target=sample)
    t.start()
    rate = capi.int_peak_sustained(sec)
    stop[0] = True
    t.join()
    print("sustained %4.2f s  %.3f T MAC32/s = %.3f of burst   sm clock min/median/max %d/%d/%d MHz, power max %.0f W"
          % (sec, rate / 1e12, rate / burst, min(clk), sorted(clk)[len(clk) // 2], max(clk), max(pw)),
          flush=True)
