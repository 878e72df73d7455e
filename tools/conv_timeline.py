"""Where does a TMA-fed convolution launch spend its time?  Per-CTA clock stamps (gb_debug_timeline) of
igemm_tma_kernel on the residual-block layer, plus launches with one agent switched off (knob 30: 1 = no loads,
2 = no MMAs, 4 = no epilogue).  Bring-up measurement, not a bench value.

    python tools/conv_timeline.py [batch]
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
    layers = [
        mb.layer("res 3x3 256->256 64x64 (+border 1)", 256, 256, 3, 1, 0, 64, 64, B, border=1),
        mb.layer("down 3x3 s2 128->256 128->64", 128, 256, 3, 2, 1, 128, 128, B),
        mb.layer("up convT 3x3 s2 256->128 64->128", 256, 128, 3, 2, 1, 64, 64, B, transposed=True, op_pad=1),
    ]
    cap = 1 << 16
    ts = torch.zeros(cap * 16, dtype=torch.int64, device="cuda")
    for L in layers:
        for what in ("fwd", "dgrad"):
            fn = mb.run(L, what)
            print(f"=== {L['name']} {what}  ({L['flops'] / 1e9:.2f} GFLOP)")
            for label, knobs in (("per-tap kernel", {9: 1}), ("per-thread store epilogue", {9: 1, 29: 1}),
                                 ("bulk-store epilogue", {9: 1, 29: 2}), ("coalesced-store epilogue", {9: 1, 29: 3}),
                                 ("no loads", {9: 1, 30: 1}), ("no MMAs", {9: 1, 30: 2}),
                                 ("no epilogue", {9: 1, 30: 4}), ("no loads, no epilogue", {9: 1, 30: 5}),
                                 ("no stats epilogue", {9: 1, 31: 1})):
                if knobs.get(31):
                    mb.NO_STATS = True
                    fn2 = mb.run(L, what)
                else:
                    mb.NO_STATS = False
                    fn2 = fn
                kk = {k: v for k, v in knobs.items() if k != 31}
                old = {k: lib.gb_debug_knob(k, v) for k, v in kk.items()}
                try:
                    t = mb.time_us(fn2, reps=10)
                    print(f"  {label:24s} {t:7.1f} us  {L['flops'] / t / 1e6:6.0f} TF", flush=True)
                finally:
                    for k, v in old.items():
                        lib.gb_debug_knob(k, v)
            for tl_label, tl_knobs in (("coalesced-store epilogue", {9: 1, 29: 3}), ("bulk-store epilogue", {9: 1, 29: 2}),
                                       ("per-thread store epilogue", {9: 1, 29: 1})):
                # time stamps of one ordinary launch (per-tap kernel, L2 warm from the launches above)
                old = {k: lib.gb_debug_knob(k, v) for k, v in tl_knobs.items()}
                ts.zero_()
                lib.gb_debug_timeline(ts.data_ptr(), cap)
                fn()
                torch.cuda.synchronize()
                lib.gb_debug_timeline(None, 0)
                for k, v in old.items():
                    lib.gb_debug_knob(k, v)
                print(f"  -- {tl_label} (epilogue mode read back: {lib.gb_debug_knob(31, 0)})")
                t = ts.view(-1, 16).cpu()
                t = t[t[:, 2] != 0]
                if len(t) == 0:
                    print("  (no stamps: launch not served by igemm_tma_kernel)")
                    continue
                g0 = int(t[:, 1].min())
                d = lambda a, b: (t[:, b] - t[:, a]).float()
                print(f"  {len(t)} CTAs on {len(set(t[:, 0].tolist()))} SMs; clock cycles, mean [min, max]:")
                for name, a, b in (("setup (barriers, TMEM alloc, taps)", 2, 3), ("first operands land", 3, 4),
                                   ("main loop (first data -> last MMA issued)", 4, 5), ("last MMA issued -> accumulator done", 5, 6),
                                   ("epilogue", 6, 7), ("  TMEM -> staged tile", 6, 8), ("  barrier", 8, 9),
                                   ("  write-out issued", 9, 10), ("  statistics", 10, 11), ("  tail (bulk-store wait, barrier)", 11, 7),
                                   ("whole CTA", 2, 7)):
                    if (t[:, a] == 0).all() or (t[:, b] == 0).all():
                        continue
                    x = d(a, b)
                    print(f"    {name:44s} {x.mean():9.0f} [{x.min():7.0f}, {x.max():7.0f}]")
                start = (t[:, 1] - g0).float() / 1e3
                print(f"    CTA start times (globaltimer): median {start.median():.1f} us, max {start.max():.1f} us; "
                      f"CTAs starting after 5 us: {(start > 5).sum().item()}")


if __name__ == "__main__":
    main()
