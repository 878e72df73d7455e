"""Per-CTA clock stamps of the persistent convolution kernel (igemm_pers_kernel, knob 16 = 3) on the residual-block
layer: does the main loop of item i + 1 keep the MMA floor while the epilogue of item i runs?  Bring-up measurement.

    python tools/pers_timeline.py [batch]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ganslate_b200 import _cabi
import conv_microbench as mb


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    lib = _cabi.lib()
    L = mb.layer("res 3x3 256->256 64x64 (+border 1)", 256, 256, 3, 1, 0, 64, 64, B, border=1)
    cap = 1 << 12
    ts = torch.zeros(cap * 16, dtype=torch.int64, device="cuda")
    for what in ("fwd", "dgrad"):
        for stats in (True, False):
            if what == "dgrad" and not stats:
                continue
            mb.NO_STATS = not stats
            fn = mb.run(L, what)
            old = {k: lib.gb_debug_knob(k, v) for k, v in {9: 1, 16: 3}.items()}
            t_us = mb.time_us(fn, reps=10)
            ts.zero_()
            lib.gb_debug_timeline(ts.data_ptr(), cap)
            fn()
            torch.cuda.synchronize()
            lib.gb_debug_timeline(None, 0)
            for k, v in old.items():
                lib.gb_debug_knob(k, v)
            t = ts.view(-1, 16).cpu()
            t = t[t[:, 2] != 0]
            print(f"=== {L['name']} {what} stats={stats}: {t_us:.1f} us, {len(t)} CTAs; cycles mean [min, max]")
            two = t[t[:, 7] != 0]     # CTAs that processed two items
            one = t[t[:, 7] == 0]
            print(f"    CTAs with two items: {len(two)}, with one: {len(one)}")
            g0 = int(t[:, 1].min())
            start, end = (t[:, 1] - g0).float() / 1e3, (t[:, 13] - g0).float() / 1e3
            wall = (t[:, 13] - t[:, 1]).float() / 1e3
            cyc = (t[:, 12] - t[:, 2]).float()
            print(f"    globaltimer: first CTA start 0, last CTA start {start.max():.1f} us, last CTA end {end.max():.1f} us; "
                  f"per-CTA wall {wall.mean():.1f} us for {cyc.mean():.0f} cycles = {cyc.mean() / wall.mean() / 1e3:.2f} GHz")
            d = lambda tt, a, b: (tt[:, b] - tt[:, a]).float()
            for name, tt, a, b in (("setup -> first operands of item 0", t, 3, 4), ("main loop item 0 (first data -> last MMA issued)", t, 4, 5),
                                   ("main loop item 1", two, 6, 7), ("item 0: last MMA issued -> accumulator seen by the epilogue", t, 5, 8),
                                   ("epilogue item 0", t, 8, 9), ("epilogue item 1", two, 10, 11),
                                   ("item 1: last MMA issued -> accumulator seen", two, 7, 10),
                                   ("whole CTA (two items)", two, 2, 12), ("whole CTA (one item)", one, 2, 12)):
                if len(tt) == 0:
                    continue
                x = d(tt, a, b)
                print(f"    {name:58s} {x.mean():9.0f} [{x.min():7.0f}, {x.max():7.0f}]")


if __name__ == "__main__":
    main()
