#!/usr/bin/env python
"""bench.py -- env-steps/sec of the batched simulator on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload invmgmt|newsvendor|netinv]

Headline workload (BASELINE.json configs[2], the env the north-star target is quoted on):
InvManagementLostSalesEnv defaults (4-stage serial chain, periods=30, Poisson mu=20), 2^24 instances per GPU,
fused 30-period rollout with the on-device base-stock policy (SF=1.0), Philox seed 5000.  One bench "step" is one
fused rollout of the whole batch = 2^24 * 30 env-steps per GPU (+ the 64-byte NCCL allreduce of the episode
statistics when N > 1).  Prints ONE JSON line (see the task contract); extra keys:
  roofline           dominant kernel (inv_jit_rollout_bs, the rollout kernel specialised for the config at run time;
                     inv_rollout_kernel when NVRTC is unavailable): issue-slot roofline (the kernel keeps its state on chip,
                     so HBM traffic is ~1 B/env-step by construction) + its HBM figures
  roofline_step_api  the HBM-bound one-period kernel (inv_step_kernel) at the same batch size, 466 B/env-step
  cpu_baseline       the C oracle port of the reference's evaluation loop on all host cores (bounded sample)
  e2e                the same rollout through the public Python API with per-episode results copied to pinned host
                     memory every step
  e2e_step_api       the host-driven policy loop: actions H2D, observation/reward/flags D2H every period (2^20 instances)
`--impl reference` times the CPU port itself (the reference is pure Python and cannot travel to the GPU box; its
in-container rates are recorded in DESIGN.md).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES_STEP_API = {"invmgmt": 466, "newsvendor": 222, "netinv": 1598}  # SURVEY.md §8d, per env-step
SM_COUNT, SCHED_PER_SM = 148, 4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="invmgmt", choices=["invmgmt", "newsvendor", "netinv"])
    ap.add_argument("--envs-per-gpu", type=int, default=0, help="override the instance count (default per workload)")
    ap.add_argument("--no-extras", action="store_true", help="skip step-API roofline, e2e and cpu baseline legs")
    return ap.parse_args()


WORKLOADS = {
    "invmgmt": dict(name="InvManagementLostSalesEnv defaults, fused 30-period rollout, on-device base-stock SF=1.0",
                    envs=1 << 24, seed=5000, periods=30),
    "newsvendor": dict(name="NewsvendorEnv defaults (lead_time=5, step_limit=40), fused rollout, classic-newsvendor policy",
                       envs=1 << 20, seed=2000, periods=40),
    "netinv": dict(name="NetInvMgmtBacklogEnv default 9-node network, fused 30-period rollout, constant-order 10% policy",
                   envs=1 << 22, seed=6000, periods=30),
}


KERNEL_SOURCES = {"invmgmt": ("invmgmt.cu", "invmgmt_jit.cu", "invmgmt_jit_args.cuh", "common.cuh", "device_rng.cuh"),
                  "newsvendor": ("newsvendor.cu", "poisson_mu.cuh", "common.cuh", "device_rng.cuh"),
                  "netinv": ("netinv.cu", "netinv.cuh", "netinv_jit.cu", "netinv_args.cuh", "common.cuh", "device_rng.cuh")}


def kernel_source_hash(workload):
    """sha1 over the CUDA sources of one env family: ties profiles/inst_counts.json to the kernels it was captured from."""
    import hashlib
    h = hashlib.sha1()
    for f in KERNEL_SOURCES[workload]:
        h.update(open(os.path.join(ROOT, "or-gym-inventory_b200", "csrc", f), "rb").read())
    return h.hexdigest()


# ---------------------------------------------------------------------------------------------------------
# clocks during the timed region
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "20"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    mx.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:  # noqa: BLE001
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------------------
# CPU baseline (C oracle port of the reference's evaluation loop), all host threads
# ---------------------------------------------------------------------------------------------------------
def cpu_port_rate(workload, seconds=10.0, threads=None):
    """env-steps/s of the oracle port on `threads` host threads over a bounded sample of about `seconds`."""
    import or_gym_inventory_b200 as pkg
    from oracle import oracle
    threads = threads or os.cpu_count() or 1

    def run(episodes):
        t0 = time.perf_counter()
        if workload == "invmgmt":
            steps, _ = oracle.invmgmt_bench(pkg.InvManagementParams(backlog=False), "base_stock", episodes, threads, seed0=5000)
        elif workload == "newsvendor":
            steps, _ = oracle.newsvendor_bench(pkg.NewsvendorParams(), "classic", episodes, threads, seed0=2000)
        else:
            P = pkg.NetInvMgmtParams()
            import numpy as np
            a = (P.spaces()[1].high * 0.1).astype(np.float32)
            steps, _ = oracle.netinv_bench(P, a, episodes, threads, seed0=6000)
        return steps, time.perf_counter() - t0

    steps, dt = run(2000 * threads)          # calibration
    episodes = max(2000 * threads, int(2000 * threads * seconds / max(dt, 1e-3)))
    steps, dt = run(episodes)
    return dict(value=steps / dt, unit="env-steps/s", cores=threads, kind="port",
                sample=f"{episodes} episodes ({steps} env-steps) of the same env/policy, reference-style PCG64+PTRS "
                       f"Poisson demand, {dt:.1f} s on {threads} threads")


def main_reference(args, emit):
    """--impl reference: the CPU implementation of the path on the box's host cores (oracle port; the Python
    reference itself cannot travel -- /root/reference does not exist on the GPU box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    W = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    import or_gym_inventory_b200 as pkg  # noqa: F401
    from oracle import oracle  # noqa: F401
    per_step_s = 1.0
    r = cpu_port_rate(args.workload, seconds=per_step_s, threads=threads)     # warm caches, calibrate
    eps_per_step = max(1000, int(r["value"] * per_step_s / W["periods"]))
    import numpy as np

    def one(seed0):
        if args.workload == "invmgmt":
            return oracle.invmgmt_bench(pkg.InvManagementParams(backlog=False), "base_stock", eps_per_step, threads, seed0=seed0)[0]
        if args.workload == "newsvendor":
            return oracle.newsvendor_bench(pkg.NewsvendorParams(), "classic", eps_per_step, threads, seed0=seed0)[0]
        P = pkg.NetInvMgmtParams()
        return oracle.netinv_bench(P, (P.spaces()[1].high * 0.1).astype(np.float32), eps_per_step, threads, seed0=seed0)[0]

    for w in range(args.warmup):
        one(W["seed"] + w)
    t0 = time.perf_counter()
    steps = 0
    for k in range(args.steps):
        steps += one(W["seed"] + 1000 + k * eps_per_step)
    dt = time.perf_counter() - t0
    val = steps / dt
    line = {"impl": "reference", "metric": "env-steps/sec", "value": val, "unit": "env-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64+f64",
            "data": "synthetic", "config": {"workload": W["name"], "sample_episodes_per_step": eps_per_step},
            "cpu_baseline": {"value": val, "unit": "env-steps/s", "cores": threads, "kind": "port",
                             "sample": f"{eps_per_step} episodes per step x {args.steps} steps, all {threads} host threads"},
            "e2e": {"value": val, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ---------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    # keep stdout clean for the ONE JSON line: libraries (e.g. NCCL's version banner) write to fd 1
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    if args.impl == "reference":
        return main_reference(args, emit)
    import torch
    import torch.distributed as dist
    import or_gym_inventory_b200 as pkg

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    host_cores = 0
    if world > 1:      # each rank streams results to pinned host memory: keep them on the GPU's own NUMA node
        from or_gym_inventory_b200.sharding import bind_host_to_gpu
        host_cores = bind_host_to_gpu(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    W = WORKLOADS[args.workload]
    N = args.envs_per_gpu or W["envs"]
    T = W["periods"]
    offset = rank * N

    if args.workload == "invmgmt":
        env = pkg.InvManagementLostSalesEnv(num_envs=N, device=dev, env_offset=offset)
        roll = lambda ep, want: env.rollout("base_stock", seed=W["seed"], episode=ep, safety_factor=1.0, want=want)  # noqa: E731
        kernel, dtype = "inv_jit_rollout_bs | inv_rollout_kernel<3,true,int>", "int32 state + f64 reward"
    elif args.workload == "newsvendor":
        env = pkg.NewsvendorEnv(num_envs=N, device=dev, env_offset=offset)
        roll = lambda ep, want: env.rollout("classic", seed=W["seed"], episode=ep, want=want)  # noqa: E731
        kernel, dtype = "nv_rollout_kernel", "f32 state + mixed f32/f64 reward"
    else:
        env = pkg.NetInvMgmtBacklogEnv(num_envs=N, device=dev, env_offset=offset)
        roll = lambda ep, want: env.rollout("constant", order_fraction=0.1, seed=W["seed"], episode=ep, want=want)  # noqa: E731
        kernel, dtype = ("net_jit_rollout" if env.specialised else "net_sim_kernel<128> (generic)"), "f64"
    want = ("ep_return", "stats", "summary")

    def step(ep):
        out = roll(ep, want)
        if world > 1:
            dist.all_reduce(out["summary"])      # the only collective: 8 float64 episode statistics
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(max(args.warmup, 3)):
        step(w)
    barrier()
    if args.workload == "invmgmt":
        kernel = ("inv_jit_rollout_bs (specialised at run time, NVRTC)" if env.rollout_specialised
                  else "inv_rollout_kernel<3,true,int>")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        out = step(100 + k)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    steps_total = float(N) * T * args.steps * world
    value = steps_total / (ms * 1e-3)
    summ = out["summary"].cpu().numpy()

    line = {"metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": {"workload": W["name"], "instances_per_gpu": N, "periods": T, "seed": W["seed"],
                       "env_steps_per_bench_step": N * T * world,
                       "host_cores_per_rank": host_cores or None,
                       "l2": "no input tensors (on-device policy + Philox demand); the per-episode outputs written "
                             f"every step ({N * 40 / 1e6:.0f} MB) exceed the 126 MB L2"},
            "gpu_launches": 2 * args.steps,
            "episode_stats": {"episodes": summ[0], "mean_return": summ[1] / max(summ[0], 1),
                              "service_level": summ[3] / max(summ[4], 1)}}
    if clocks is not None:
        line["clocks"] = {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"],
                          "samples": clocks["samples"]}

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    # ---- e2e (every rank): the public evaluate() API -- per-episode results land in pinned host memory every step;
    # the copy of step k overlaps the kernel of step k+1 (two buffer sets, second stream); host waits for every result
    if not args.no_extras:
        pol = {"invmgmt": ("base_stock", dict(safety_factor=1.0)), "newsvendor": ("classic", {}),
               "netinv": ("constant", dict(order_fraction=0.1))}[args.workload]
        reps = max(4, min(args.steps, 20))

        def run_e2e(e2e_want, first):
            d2h = 0
            for res in env.evaluate(pol[0], episodes=3, seed=W["seed"], first_episode=first, want=e2e_want, **pol[1]):
                d2h = sum(v.numel() * v.element_size() for v in res.values())
            barrier()
            t0 = time.perf_counter()
            chk = 0.0
            for res in env.evaluate(pol[0], episodes=reps, seed=W["seed"], first_episode=first + 100, want=e2e_want, **pol[1]):
                chk += float(res["summary"][0]) + float(res["ep_return"][-1]) * 0.0      # touch the host copies
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            assert chk == float(N) * reps
            return float(N) * T * reps * world / float(dt.item()), d2h * world

        # headline: what the reference's evaluation report needs -- every episode's return (histograms, quantiles) and
        # the 8 aggregate statistics (mean / std of return, service level, stock-outs, inventory: process_results)
        v1, b1 = run_e2e(("ep_return", "summary"), 300)
        # everything the reference keeps per episode (return + sales / demand / stock-out / inventory sums): PCIe-bound
        full_want = ("ep_return", "stats32", "summary") if args.workload == "invmgmt" else want
        v2, b2 = run_e2e(full_want, 600)
        # only the 8 aggregate statistics (what the reference's benchmark scripts finally print): kernel-bound
        def run_summary_only():
            for res in env.evaluate(pol[0], episodes=3, seed=W["seed"], first_episode=900, want=("summary",), **pol[1]):
                pass
            barrier()
            t0 = time.perf_counter()
            chk = 0.0
            for res in env.evaluate(pol[0], episodes=reps, seed=W["seed"], first_episode=1000, want=("summary",), **pol[1]):
                chk += float(res["summary"][0])
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            assert chk == float(N) * reps
            return float(N) * T * reps * world / float(dt.item())
        v3 = run_summary_only()
        line["e2e"] = {"value": v1, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": b1, "steps": reps,
                       "summary_only": {"value": v3, "d2h_bytes_per_step": 64 * world,
                                        "note": "only the 8 aggregate statistics cross PCIe (kernel-bound)"},
                       "note": "public API env.evaluate(): this workload's inputs are the policy/seed scalars passed as "
                               "kernel parameters (no input tensors); every episode's float64 return and the 8 aggregate "
                               "statistics are copied to pinned host memory each step and consumed by the host inside the "
                               "timed region; copy of step k overlaps the kernel of step k+1",
                       "with_per_episode_statistics": {"value": v2, "d2h_bytes_per_step": b2,
                                                       "note": "additionally the four per-episode statistics of the "
                                                               "reference's evaluate_agent rows (PCIe-bound)"}}
    else:
        line["e2e"] = None

    if rank == 0 and not args.no_extras:
        # ---- roofline of the dominant kernel (fused rollout): issue-slot bound ---------------------------------
        counts = {}
        try:
            counts = json.load(open(os.path.join(ROOT, "profiles", "inst_counts.json")))
        except Exception:  # noqa: BLE001
            pass
        for _ in range(3):
            roll(0, want)
        torch.cuda.synchronize()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        reps = max(5, args.steps)
        for k in range(reps):
            roll(200 + k, want)
        k1.record()
        torch.cuda.synchronize()
        kms = k0.elapsed_time(k1) / reps
        alg_bytes = N * 40.0           # per launch: 8 B return + 32 B statistics per episode
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        issue_peak = SM_COUNT * SCHED_PER_SM * sm_mhz * 1e6           # warp-instructions / s at the observed clock
        ckey = args.workload
        if args.workload == "invmgmt" and "inv_jit" not in kernel:
            ckey = "invmgmt_aot"       # NVRTC unavailable: the ahead-of-time kernel ran (its own instruction count)
        c = counts.get(ckey, {})
        src_now = kernel_source_hash(args.workload)
        stale = bool(c) and c.get("src_sha1") not in (None, src_now)
        roof = {"kernel": kernel, "bound": "issue", "unit": "warp-inst/s", "peak": issue_peak,
                "peak_source": f"148 SMs x 4 schedulers x {sm_mhz:.0f} MHz (median SM clock sampled during the timed region)",
                "kernel_ms": kms, "achieved": None, "frac": None,
                "traffic": (c["dram_bytes_per_env_step"] * N * T) if c.get("dram_bytes_per_env_step") else None,
                "hbm": {"achieved_gbs": alg_bytes / (kms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                        "frac": alg_bytes / (kms * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes_per_launch": alg_bytes,
                        "peak_source": peak_src}}
        if stale:
            roof["note"] = ("instruction count in profiles/inst_counts.json was captured for different kernel sources "
                            "(sha1 mismatch): issue-roofline fraction withheld until the capture is redone")
        if c.get("warp_inst_per_env_step") and not stale:
            inst = c["warp_inst_per_env_step"] * N * T
            roof["achieved"] = inst / (kms * 1e-3)
            roof["frac"] = roof["achieved"] / issue_peak
            roof["warp_inst_per_launch"] = inst
            roof["inst_source"] = c.get("source")
        line["roofline"] = roof

        # ---- HBM-bound one-period kernel (step API) at the same batch size ---------------------------------------
        if args.workload == "invmgmt":
            a = torch.randint(0, 100, (N, 3), dtype=torch.int64, device=dev)
            sk = "inv_step_kernel<3,true,int>"
        elif args.workload == "newsvendor":
            a = torch.rand((N, 1), device=dev) * 100
            sk = "nv_step_kernel"
        else:
            a = torch.rand((N, len(env.reorder_links)), device=dev) * 100
            sk = "net_jit_step" if env.specialised else "net_sim_kernel<128> (generic, STEP)"
        env.reset(seed=W["seed"])
        for _ in range(3):
            env.step(a)
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(20):
            env.step(a)
        s1.record()
        torch.cuda.synchronize()
        sms = s0.elapsed_time(s1) / 20
        ab = ALG_BYTES_STEP_API[args.workload]
        # optional info tensors written on top of the algorithmic bytes (Python host default, info_level=1)
        if args.workload == "invmgmt":
            ib = 8 + 8 * 4 + 8 * 4 + 8                      # demand, sales[m], unfulfilled[m], period profit
        elif args.workload == "newsvendor":
            ib = 8 + 8 * 4                                  # demand, four reward parts
        else:
            nE, nM, nJ = len(env.reorder_links), len(env.retail_links), len(env.main_nodes)
            ib = 8 * (nM + (nE + nM) + nJ + 1)              # demand, sales, node profit, period profit
        # the same kernel without the optional info tensors = exactly the algorithmic bytes of SURVEY.md §8d
        env.close()
        del env
        torch.cuda.empty_cache()
        cls0 = {"invmgmt": pkg.InvManagementLostSalesEnv, "newsvendor": pkg.NewsvendorEnv,
                "netinv": pkg.NetInvMgmtBacklogEnv}[args.workload]
        env0 = cls0(num_envs=N, device=dev, env_offset=offset, info_level=0)
        env0.reset(seed=W["seed"])
        for _ in range(3):
            env0.step(a)
        torch.cuda.synchronize()
        s0.record()
        for _ in range(20):
            env0.step(a)
        s1.record()
        torch.cuda.synchronize()
        sms0 = s0.elapsed_time(s1) / 20
        ach = N * ab / (sms0 * 1e-3) / 1e9
        ach_i = N * (ab + ib) / (sms * 1e-3) / 1e9
        line["roofline_step_api"] = {"kernel": sk, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                                     "frac": ach / hbm_peak,
                                     "traffic": (counts.get(args.workload + "_step", {}).get("dram_bytes_per_env_step") or 0) * N or None,
                                     "kernel_ms": sms0, "env_steps_per_s": N / (sms0 * 1e-3),
                                     "algorithmic_bytes_per_env_step": ab, "instances": N, "peak_source": peak_src,
                                     "note": "env built with info_level=0: the kernel moves exactly the algorithmic bytes "
                                             "(actions, observation, reward, flags, state read+write); `traffic` is the "
                                             "ncu DRAM byte count of the capture WITH info tensors",
                                     "with_info_tensors": {"kernel_ms": sms, "bytes_per_env_step": ab + ib,
                                                           "achieved": ach_i, "frac": ach_i / hbm_peak,
                                                           "env_steps_per_s": N / (sms * 1e-3),
                                                           "frac_counting_algorithmic_bytes_only": N * ab / (sms * 1e-3) / 1e9 / hbm_peak}}
        env0.close()
        del a, env0
        # ---- the host-driven loop (a policy on the CPU: numpy actions in, observations out every period) ------------
        # Same step kernel, but the Gymnasium tensors cross PCIe both ways each period and the host waits for them:
        # this is what the reference's agent loop costs when only the env moves to the GPU.
        Ns = min(N, 1 << 20)
        envh = cls0(num_envs=Ns, device=dev, env_offset=offset, info_level=0)
        obs_d, _ = envh.reset(seed=W["seed"])
        if args.workload == "invmgmt":
            a_h = torch.randint(0, 100, (Ns, 3), dtype=torch.int64).pin_memory()
        elif args.workload == "newsvendor":
            a_h = (torch.rand((Ns, 1)) * 100).pin_memory()
        else:
            a_h = (torch.rand((Ns, len(envh.reorder_links))) * 100).pin_memory()
        a_d = torch.empty_like(a_h, device=dev)
        obs_h = torch.empty(obs_d.shape, dtype=obs_d.dtype).pin_memory()
        rew_h = torch.empty(Ns, dtype=torch.float64).pin_memory()
        tr_h = torch.empty(Ns, dtype=torch.uint8).pin_memory()

        def host_step():
            a_d.copy_(a_h, non_blocking=True)
            o, r, _, tr, _ = envh.step(a_d)
            obs_h.copy_(o, non_blocking=True)
            rew_h.copy_(r.view(-1), non_blocking=True)
            tr_h.copy_(tr.view(-1).view(torch.uint8), non_blocking=True)
            torch.cuda.synchronize()
            return float(rew_h[0]) + float(obs_h.view(-1)[0])      # the host policy reads what came back

        for _ in range(3):
            host_step()
        hs = 10
        t0 = time.perf_counter()
        for _ in range(hs):
            host_step()
        hdt = time.perf_counter() - t0
        line["e2e_step_api"] = {"value": Ns * hs / hdt, "unit": "env-steps/s", "instances": Ns, "steps": hs,
                                "h2d_bytes_per_step": a_h.numel() * a_h.element_size(),
                                "d2h_bytes_per_step": obs_h.numel() * obs_h.element_size() + Ns * 9,
                                "note": "host-driven policy loop on rank 0: actions from pinned host memory, observation + "
                                        "reward + truncated flags back to pinned host memory, host waits every period "
                                        "(PCIe-bound; the fused on-device policies exist to avoid exactly this)"}
        envh.close()
        del envh, a_d, a_h, obs_h
        torch.cuda.empty_cache()
        if world == 1:
            line["cpu_baseline"] = cpu_port_rate(args.workload, seconds=10.0)
            # the other BASELINE.json configs, measured briefly with the same method (fused rollout, CUDA events)
            others = {}
            for wname in ("newsvendor", "netinv"):
                if wname == args.workload:
                    continue
                Wo = WORKLOADS[wname]
                No, To = Wo["envs"], Wo["periods"]
                if wname == "newsvendor":
                    eo = pkg.NewsvendorEnv(num_envs=No, device=dev)
                    ro = lambda ep: eo.rollout("classic", seed=Wo["seed"], episode=ep)  # noqa: E731
                else:
                    eo = pkg.NetInvMgmtBacklogEnv(num_envs=No, device=dev)
                    ro = lambda ep: eo.rollout("constant", order_fraction=0.1, seed=Wo["seed"], episode=ep)  # noqa: E731
                for k in range(3):
                    ro(k)
                torch.cuda.synchronize()
                o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                o0.record()
                for k in range(20):
                    ro(10 + k)
                o1.record()
                torch.cuda.synchronize()
                oms = o0.elapsed_time(o1) / 20
                others[wname] = {"workload": Wo["name"], "instances": No, "periods": To, "ms_per_rollout": oms,
                                 "env_steps_per_s": No * To / (oms * 1e-3)}
                eo.close()
                del eo
                torch.cuda.empty_cache()
            if args.workload == "invmgmt":
                # the north-star target is quoted on InvManagementBacklogEnv; the random policy is cfg 3's second driver
                for label, cls_o, pol_o, kw_o in (
                        ("invmgmt_backlog_base_stock", pkg.InvManagementBacklogEnv, "base_stock", dict(safety_factor=1.0)),
                        ("invmgmt_lost_sales_random", pkg.InvManagementLostSalesEnv, "random", {})):
                    eo = cls_o(num_envs=N, device=dev)
                    for k in range(3):
                        eo.rollout(pol_o, seed=W["seed"], episode=k, **kw_o)
                    torch.cuda.synchronize()
                    o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    o0.record()
                    for k in range(20):
                        eo.rollout(pol_o, seed=W["seed"], episode=10 + k, **kw_o)
                    o1.record()
                    torch.cuda.synchronize()
                    oms = o0.elapsed_time(o1) / 20
                    others[label] = {"workload": f"{cls_o.__name__} defaults, fused rollout, {pol_o} policy", "instances": N,
                                     "periods": T, "ms_per_rollout": oms, "env_steps_per_s": N * T / (oms * 1e-3),
                                     "specialised_kernel": eo.rollout_specialised}
                    eo.close()
                    del eo
                    torch.cuda.empty_cache()
            line["other_configs"] = others
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
