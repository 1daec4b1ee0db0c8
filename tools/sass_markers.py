#!/usr/bin/env python
"""tools/sass_markers.py -- regenerate profiles/*_sass_* from force2vec_b200/lib/libf2v.so:
per-kernel counts of the SASS mnemonics that show what the kernels use (TMA bulk copies, mbarrier,
float4 gathers, packed fp32x2 math, programmatic dependent launch, system-scope flag traffic) and
do not use (tensor cores), plus full listings of the default kernels."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "force2vec_b200", "lib", "libf2v.so")
OUT = os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r1"

PAT = [("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("LDG.E.128", r"LDG\.E\.128"), ("LDG total", r"\bLDG"),
       ("STG.E.128", r"STG\.E\.128"), ("SHFL", r"\bSHFL"), ("FFMA2", r"\bFFMA2"), ("FFMA", r"\bFFMA\b"),
       ("DADD+DMUL", r"\b(DADD|DMUL)"), ("ATOMG", r"\bATOMG"), ("ACQBULK/PDL (griddepcontrol)", r"ACQBULK|PREEXIT|DEPBAR\.LE SB0, 0x0 ;.*griddep"),
       ("LDG.STRONG.SYS", r"LDG\.E\.64\.STRONG\.SYS"), ("STG.STRONG.SYS", r"STG\.E\.64\.STRONG\.SYS"),
       ("MEMBAR.SYS", r"MEMBAR\.\w+\.SYS"), ("LDGSTS (cp.async)", r"\bLDGSTS"), ("LDS.128", r"LDS\.128"), ("FMNMX", r"\bFMNMX"),
       ("VOTE", r"\bVOTE"), ("multimem (MULTIMEM/ST...MMIO)", r"MULTIMEM|\.MMIO"), ("UTC*MMA", r"UTC\w*MMA"), ("HMMA", r"\bHMMA")]


def main():
    sass = subprocess.check_output(["cuobjdump", "-sass", LIB]).decode()
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    lines = ["SASS markers per kernel of force2vec_b200/lib/libf2v.so (cuobjdump -sass), %s" % TAG,
             "UBLKCP = cp.async.bulk (TMA bulk copy g->s); SYNCS = mbarrier ops; LDG.E.128 = float4 gathers; "
             "FFMA2 = packed fp32x2; *.STRONG.SYS / MEMBAR.SYS = multi-GPU exchange flags; "
             "no UTC*MMA/HMMA (no tensor cores by design)", ""]
    keep = {}
    for f in funcs:
        name = f.split("\n", 1)[0].strip()
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        body = [l for l in f.split("\n") if re.search(r"/\*[0-9a-f]{4,6}\*/", l)]
        text = "\n".join(body)
        counts = ["instructions %d" % len(body)] + ["%s %d" % (k, len(re.findall(p, text))) for k, p in PAT]
        lines += [dem, "  " + " | ".join(counts)]
        keep[dem] = f
    open(os.path.join(OUT, "%s_sass_markers.txt" % TAG), "w").write("\n".join(lines) + "\n")
    want = {"force_batch_d128_opt5_5cta": "force_batch_kernel<f2v::VecL<128, 16, 2, 5>, 5>",       # cfg4, the headline kernel
            "force_batch_d128_opt6_5cta": "force_batch_kernel<f2v::VecL<128, 16, 2, 5>, 6>",       # cfg2
            "force_batch_d64_opt7_5cta": "force_batch_kernel<f2v::VecL<64, 8, 2, 5>, 7>",          # cfg3
            "force_batch_d128_opt5_ring_s4": "force_batch_kernel<f2v::RingL<128, 16, 4, 3>, 5>",   # asynchronous ring (LDGSTS)
            "force_flow_d128_opt6": "force_flow_kernel<f2v::VecL<128, 16, 8, 2>, 6>",              # dataflow epoch
            "walk_kernel": "walk_kernel(", "bcast_rows": "bcast_rows_kernel<true>", "peer_sync": "peer_sync_kernel(",
            "checksum": "checksum_kernel("}
    for tag, needle in want.items():
        for dem, f in keep.items():
            if needle in dem:
                body = "\n".join(l for l in f.split("\n") if not re.match(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", l))
                open(os.path.join(OUT, "%s_sass_%s.sass" % (TAG, tag)), "w").write("Function : " + body)
                break
    print("wrote markers for %d kernels" % len(keep))


if __name__ == "__main__":
    main()
