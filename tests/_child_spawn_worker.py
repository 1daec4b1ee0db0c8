"""Worker for tests/test_dist_cpu.py::test_bench_children_*: run under torch.distributed.run (gloo, CPU) it plays
the parent of bench.py's N = 8 extras (bench.cfg5_extras); started by bench.run_child with --child it plays the
child: a process with this rank's RANK / WORLD_SIZE that must be able to form its OWN process group."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist


def child(argv):
    mode = argv[argv.index("--child") + 1]
    if mode == "crash":
        os.abort()
    if mode == "hang":
        time.sleep(600)
    assert "TORCHELASTIC_USE_AGENT_STORE" not in os.environ
    dist.init_process_group("gloo")
    t = torch.tensor([dist.get_rank() + 1.0])
    dist.all_reduce(t)
    if mode == "rank1_fails" and dist.get_rank() == 1:
        sys.exit(7)
    if dist.get_rank() == 0:
        w = dist.get_world_size()
        print("noise on stdout")
        print(json.dumps({"value": float(t.item()), "ms_per_step": 1.5, "n_gpus": w, "steps": 3, "warmup": 2,
                          "config": {"workload": "child " + mode, "n": 4, "nnz": 6}, "parity": {"bit_exact": True},
                          "argv": argv}), flush=True)
    dist.destroy_process_group()


def parent():
    import bench
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    me = os.path.abspath(__file__)
    runs = (("ok", ["--child", "ok"], 1, None), ("not_needed", ["--child", "crash"], 1, "ok"), ("crash", ["--child", "crash"], 1, None),
            ("after_crash", ["--child", "ok"], 1, "crash"), ("rank1_fails", ["--child", "rank1_fails"], 1, None))
    out = bench.cfg5_extras(dist, rank, world, runs=runs, budget_s=(time.time() - bench.T_START) + 120, script=me)
    # a child that hangs is killed when its share of the budget is over; a run that no longer fits is not started
    out.update(bench.cfg5_extras(dist, rank, world, runs=(("hang", ["--child", "hang"], 1, None),),
                                 budget_s=(time.time() - bench.T_START) + 34, script=me))
    out.update(bench.cfg5_extras(dist, rank, world, runs=(("late", ["--child", "ok"], 50, None),),
                                 budget_s=(time.time() - bench.T_START) + 10, script=me))
    dist.barrier()
    if rank == 0:
        print("PARENT_RESULT " + json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    if "--child" in sys.argv:
        child(sys.argv[1:])
    else:
        parent()
