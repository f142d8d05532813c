"""Regenerates the machine-derived parts of profiles/r2_*.md / r2_traffic.json from gpurun_out/ (run here, on the CPU box, after
tools/r2_configs.sh, tools/parity_report.py, the launch-list and `ncu --set full` captures came back):
    python tools/make_r2_profiles.py
The hand-written readings at the top of each file are kept in this script so that they travel with the numbers."""
import csv
import json
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "profiles"
G = ROOT / "gpurun_out"


def run(*a):
    return subprocess.run(list(a), capture_output=True, text=True, cwd=ROOT).stdout


def bench_rows():
    rows, hdr, eng = [], None, None
    for l in (G / "r2_configs.log").read_text().splitlines():
        if l.startswith("## engine"):
            eng = l[3:]
        elif l.startswith("##"):
            hdr = l[3:]
        elif l.startswith("{"):
            rows.append((eng, hdr, json.loads(l)))
            eng = None
    return rows


def kline(d):
    km = d["roofline"]["kernel_ms"]
    return (f"{d['value'] / 1e6:.1f} | {d['ms_per_step']:.3f} | {km['rollout_fwd']:.3f} | {km['mlp_rows']:.3f} | "
            f"{km['rollout_jac']:.3f} | {km['rollout_bwd']:.3f}")


def main():
    default = json.loads((G / "r2_bench_default.json").read_text().strip().splitlines()[-1])
    km = default["roofline"]["kernel_ms"]
    # ---- launches -------------------------------------------------------------------------------------------------------------
    ls = run("python", "tools/launch_summary.py", "gpurun_out/r2_launches.csv")
    (OUT / "r2_launches_summary.md").write_text(f"""# Round 2: launch list of one SHAC episode (ncu --metrics gpu__time_duration.sum, cold-cache serialised; compare SHARES)

command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline`
(after the same command exited 0 without ncu; full CSV: r2_launches_bench.csv; summarised by `tools/launch_summary.py`: the rows are the
launches between the last two L2-flush fills of bench.py, i.e. one timed episode of BalanceCartPole + boundary projection, 4096 envs, H = 64)

{ls}
Round 1 had 42 launches per episode, 8 of them ours (85 % of the device time).  What is left of the library: three fills (`flat.zero_()`
of the gradient buffer and two schedule fills), one `index_select` of the terminal observation rows and two integer ops of the
intervention counter.  The `step_fwd_kernel` / second `mlp_rows_tc_kernel` launch belong to the e2e leg's `observation` / row-0 value
refresh that falls inside the window.

Live CUDA-event times of `bench.py` (same binary, 100 timed episodes: {default['ms_per_step']:.3f} ms per episode = {default['value'] / 1e6:.1f} M
env-steps/s): forward {km['rollout_fwd']:.3f} ms, critic values {km['mlp_rows']:.3f}, step Jacobians {km['rollout_jac']:.3f}, reverse pass
{km['rollout_bwd']:.3f} -- the ncu shares above agree.  Device timeline of the same episode: r2_timeline.md.
""")
    tl = (G / "r2_timeline.txt").read_text()
    tl = tl[tl.index("episode span"):]
    (OUT / "r2_timeline.md").write_text(f"""# Round 2: device timeline of one graph-replayed episode (torch.profiler / CUPTI, `python tools/episode_breakdown.py --gpu-timeline`)

Third of four back-to-back episodes (BalanceCartPole + boundary projection, 4096 envs, H = 64, fp32), after 8 warm-up episodes, final
round-2 binary: ONE graph per episode, the critic values / episode scalars / loss value on a branch stream beside the reverse pass
(negative "gap before" = the launch overlaps the previous line).  The critic kernel and `rollout_jac_kernel` do not really share the
SMs: the critic's 512 threads x 128 registers fill the register file, so the Jacobian kernel's CTAs start as the critic's retire
(its 45 us stretch to ~170 us of wall time).  Before the loss value moved behind the forward kernel the same trace showed 444 us of idle
device time per 1.63 ms episode under the profiler (`update_policy`'s `float(loss)` waited for the whole reverse pass, and the host
prepared the next episode only after that); now the host prologue of episode i + 1 runs under the reverse pass of episode i.

```
{tl}```
""")
    # ---- traffic -----------------------------------------------------------------------------------------------------------------
    raw = run("ncu", "-i", "gpurun_out/r2_full.ncu-rep", "--page", "raw", "--csv")
    r = list(csv.reader(raw.splitlines()))
    hdr, units = r[0], r[1]

    def val(row, key):
        i = hdr.index(key)
        return float(row[i].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
    per = {}
    for row in r[2:]:
        name = row[hdr.index("Kernel Name")]
        b = val(row, "dram__bytes_read.sum") + val(row, "dram__bytes_write.sum")
        key = ("rollout_fwd" if "rollout_wm" in name else "mlp_rows" if "mlp_rows_tc" in name else "mlp_rows_vjp" if "vjp" in name
               else "rollout_jac" if "rollout_jac" in name else "rollout_scan" if "scan" in name
               else "policy_rows_bwd_grad" if re.search(r"<1, 5, 5, 1, 4>", name) else "policy_rows_bwd_jac")
        per[key] = int(b)
    per["rollout_bwd"] = per["policy_rows_bwd_jac"] + per["rollout_scan"] + per["policy_rows_bwd_grad"]
    json.dump({"source": "profiles/r2_ncu_summary.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch; capture "
                         "gpurun_out/r2_full.ncu-rep of the final round-2 binary; rollout_bwd = Jacobian mode + scan + gradient mode)",
               "workload": {"env": "BalanceCartPole", "safeguard": "boundary_projection", "num_envs_per_gpu": 4096, "horizon": 64, "dtype": "f32"},
               "bytes_per_launch": per}, open(OUT / "r2_traffic.json", "w"), indent=1)
    # ---- ncu summary -------------------------------------------------------------------------------------------------------------
    s4 = run("python", "tools/ncu_summary.py", "gpurun_out/r2_full.ncu-rep").replace("# ncu summary of `gpurun_out/r2_full.ncu-rep`", "")
    s64 = run("python", "tools/ncu_summary.py", "gpurun_out/r2_full_64k.ncu-rep").replace("# ncu summary of `gpurun_out/r2_full_64k.ncu-rep`", "")
    hot = run("python", "tools/ncu_hot_lines.py", "gpurun_out/r2_full.ncu-rep", "rollout_wm2", "14") + \
        run("python", "tools/ncu_hot_lines.py", "gpurun_out/r2_full.ncu-rep", "policy_rows_bwd", "12")
    (OUT / "r2_ncu_summary.md").write_text(f"""# Round 2: `ncu --set full` summaries of the whole-horizon kernels (final round-2 binary)

Captured on a B200 with `ncu --set full --clock-control none --import-source on -k regex:<kernels> -c 7 python bench.py --steps 1 --warmup 3
--no-cpu-baseline` (one GPU, after the same command exited 0 without ncu), read on the CPU box with `tools/ncu_summary.py`,
`tools/ncu_hot_lines.py` (per CUDA source line: share of stall samples / of executed warp instructions).  No number below is a
bench value: ncu serialises launches and replays each kernel; the live per-kernel times are in the bench line (`roofline.kernel_ms`).

## What the captures say

* **Forward, 4096 envs (`rollout_wm2_fwd_kernel`, 329 us under ncu, 0.35 ms live)**: 128 CTAs x 8 warps (4 pairs of warps, 8 environments
  per pair), 162 registers, no spills.  88 M warp instructions (round 1 FMA kernel: 337 M), 0.30 issue slots / cycle / SMSP, tensor pipe
  25 % (144 `HMMA.16816.F32` per warp-step at the measured 8 cycles per HMMA per sub-partition, `tools/micro/hmma_rate.cu`).  Stalls:
  `wait` 1.6 (fixed-latency dependency chains), `barrier` 1.6 (the pair's second warp idles at the step barrier while the first one runs
  the environment step: attributed by ncu to the first instruction after the barrier), `short_scoreboard` 0.9.  The environment step
  was 48 % of the stall samples of the first warp-MMA kernel; after the angle wrap lost its two range reductions + atan2, the gate its
  square root and divisions and the interval rows their second division it is 22 % (25 M of 88 M instructions).  DRAM: 5.5 MB read,
  228 MB written (the two ELU stashes, 1 KB per env-step, in the reverse pass's tile layout) = 11x the 20.9 MB of algorithmic bytes, on
  purpose (HBM is 6 % busy).
* **Reverse pass, 4096 envs**: Jacobian-mode `policy_rows_bwd_kernel` 103 us (first round-2 version: 163 us), gradient mode 238 us
  (351 us), scan 35 us (55 us before its rows streamed through a cooperative cp.async ring).  With the stash in 128-row tiles of 16-byte chunks the `lg_throttle` stall is gone (2.0 -> < 0.3) and
  `long_scoreboard` fell from 3.6 to 2.6 (gradient mode) / 2.2 to 1.6 (Jacobian mode); issue rate 0.29 -> 0.43 / 0.34 -> 0.52.  The
  gradient kernel is now bound by its instruction count (106 M warp instructions: the four column-sum transposes through shared memory are
  35 % of them) and the MMA round trip (`mbar_wait` 8 % of samples).  DRAM traffic = the stash read once per mode (273 / 281 MB).  The scan
  (one thread per environment, 64 sequential steps) is pure latency; what made it slow was not the depth of its prefetch but the shape
  of its loads: every thread read its own 240-byte row, i.e. each load instruction of a warp touched 32 different lines.  Now the warp
  fills a 4-stage shared-memory ring cooperatively (512 contiguous bytes per cp.async instruction).
* **65 536 envs** (`r2_full_64k`): tcgen05 forward 2.46 ms writing 4.5 GB of stash (1.8 TB/s), Jacobian mode 1.44 ms reading 4.36 GB
  (3.0 TB/s = 46 % of the measured HBM peak), gradient mode 3.18 ms (1.4 TB/s), critic 1.85 ms, scan 0.30 ms (5.5 TB/s of Jacobian
  rows: HBM-bound there).  At this size the row-batched kernels are the episode (reverse 47 %, critic 18 %) and the Jacobian mode is
  within 2.2x of its HBM floor; the next lever there is a 16-bit stash (halves 3 KB per env-step of traffic), not more tensor-core work.
* **Critic (`mlp_rows_tc_kernel`)**: 149 -> 129 us with the hardware-exponential ELU (78 M -> 59 M warp instructions), tensor pipe 19 %,
  still single-buffered: two 128-row activation operands do not fit beside two hi/lo weight matrices (256 KB > 227 KB), and M = 64 half
  tiles would idle half the lanes of every `tcgen05.ld` epilogue; see DESIGN.md section 8.

## Hottest source lines

```
{hot}```

## 4096 envs (bench default): per kernel
{s4}

## 65 536 envs: per kernel
{s64}
""")
    # ---- engines / configs -------------------------------------------------------------------------------------------------------
    rows = bench_rows()
    names = {"engine 1": "tcgen05 CTA tiles (SGBPO_ROLLOUT_ENGINE=1)", "engine 2": "FMA register tiles (=2)", "engine 3": "warp-MMA (=3)"}
    t = ("| forward engine | envs | M env-steps/s | ms / episode | forward | critic values | step Jacobians | reverse pass |\n"
         "|---|---|---|---|---|---|---|---|\n")
    for eng, hdr, d in rows:
        if eng:
            t += f"| {names[eng]} | {d['config']['num_envs_per_gpu']} | {kline(d)} |\n"
    t2 = ("| envs | forward engine picked | M env-steps/s | ms / episode | forward | critic values | step Jacobians | reverse pass |\n"
          "|---|---|---|---|---|---|---|---|\n")
    for eng, hdr, d in rows:
        if not eng and hdr.startswith("--num-envs"):
            t2 += f"| {d['config']['num_envs_per_gpu']} | {d['roofline']['engines']['forward']} | {kline(d)} |\n"
    (OUT / "r2_engines.md").write_text(f"""# Round 2: forward / reverse engines, measured (B200, fp32, BalanceCartPole + boundary projection, H = 64, policy [128, 128])

All lines are `python bench.py --num-envs N --steps 20 --warmup 3 --no-cpu-baseline` of the final round-2 binary (`tools/r2_configs.sh`,
raw JSON: gpurun_out/r2_configs.log), kernel columns = live CUDA-event times in ms.  Reverse pass = parallel engine in every row
(row-batched tcgen05 policy backward in Jacobian mode, per-environment scan, row-batched tcgen05 policy backward in gradient mode).

## Forward engines side by side (forced with SGBPO_ROLLOUT_ENGINE)

{t}
* The tcgen05 forward needs 128 environments per CTA: 32 CTAs at 4096 envs, 128 at 16 384; its time is flat until the GPU is full, then
  linear.  The warp-MMA forward puts 8 environments on a warp (a pair of warps up to 4736 environments) and has no block-wide barrier
  in the time loop.  Cross-over between the two: ~11 k environments -> `FusedRollout.TC_FORWARD_MIN_ENVS = 12288`.  The FMA tiles lose
  everywhere in fp32 and remain the float64 / other-width engine.
* Why mma.sync and not tcgen05 below the cross-over: `tools/micro/hmma_rate.cu` measures 8.1 cycles per `mma.sync.m16n8k16` per SM
  sub-partition (0.5 per cycle per SM, 21 cycles latency).  288 of them per 8 environments and step are 2.3 k cycles, against ~4 k issue
  cycles of FFMA for the same rows and against a shared-memory -> MMA -> commit -> mbarrier -> tcgen05.ld round trip per layer.

## Size sweep with the default engine choice

{t2}
Round 1 (FMA forward, sequential FMA reverse): 133 M env-steps/s at 4096 envs, 228 M at 65 536.

## Steps of the round on the headline workload (4096 envs; ms per episode, forward / reverse pass)

| change | ms / episode | forward | reverse |
|---|---|---|---|
| round 1: FMA forward, sequential FMA reverse + cuBLAS SIMT dW_hid GEMM | 1.98 | 0.60 | 0.71 + 0.18 |
| parallel reverse pass on tcgen05 (row-major stash, e2 rows by 512-byte TMA copies) | 1.72 | 0.60 | 0.58 |
| one Philox input kernel instead of ~20 library launches | 1.71 | 0.60 | 0.58 |
| stash in 128-row tiles of 16-byte chunks, one 64 KB TMA copy per tile, 16-byte vector reductions for dW_hid | 1.63 | 0.65 | 0.42 |
| warp-MMA forward | 1.45 | 0.48 | 0.42 |
| episodes timed back to back; loss value behind the forward kernel (host prologue under the reverse pass) | 1.22 | 0.48 | 0.42 |
| episode scalars; one graph per episode with the critic branch | 1.19 | 0.48 | 0.42 |
| pair-split warp-MMA forward | 1.16 | 0.45 | 0.42 |
| angle wrap without sin / cos / atan2 | 1.11 | 0.41 | 0.42 |
| gate without sqrt / divisions, one reciprocal per interval row | 1.05 | 0.35 | 0.42 |
| hardware-exponential ELU in the critic, L2 prefetches, one-launch clip + Adam | 1.02 | 0.35 | 0.41 |
| scan rows through a cooperative cp.async ring | 1.00 | 0.35 | 0.39 |

Column split 8 instead of 4 in the row-batched kernels (`SGBPO_ROWS_BWD_SPLIT=8`) was measured slower (gradient mode 0.416 vs 0.351 ms
with the row-major stash) and stays off.
""")
    t3 = ("| config | engine | M env-steps/s | e2e M env-steps/s | ms / episode | infeasible / guarded steps | nonfinite | gates |\n"
          "|---|---|---|---|---|---|---|---|\n")
    for eng, hdr, d in rows:
        if eng or hdr.startswith("--num-envs"):
            continue
        ep = d["gates"]["episodes"]
        extra = f" ({ep['set_exits_reference_quirk_Q2']} set exits, reference quirk Q2)" if "set_exits_reference_quirk_Q2" in ep else ""
        t3 += (f"| `{hdr}` | {d['roofline']['engines']['forward']} | {d['value'] / 1e6:.2f} | {d['e2e']['value'] / 1e6:.2f} | "
               f"{d['ms_per_step']:.3f} | {ep['infeasible']} / {ep['guard_used']} | {ep['nonfinite']} | "
               f"{'pass' if d['gates']['pass'] else 'FAIL'}{extra} |\n")
    mg = {}
    for n in (1, 2, 4, 8):
        p = G / f"mg{n}.json"
        if p.exists():
            mg[n] = json.loads(p.read_text().strip().splitlines()[-1])
    mgs = " / ".join(f"{mg[n]['value'] / 1e6:.0f}" for n in sorted(mg))
    mgt = " / ".join(f"{mg[n]['ms_per_step']:.3f}" for n in sorted(mg))
    eff = " / ".join(f"{mg[n]['value'] / (n * mg[1]['value']):.3f}" for n in sorted(mg) if n > 1)
    (OUT / "r2_configs.md").write_text(f"""# Round 2: bench lines of the other BASELINE.json configs (B200, 1 GPU unless stated, fp32, final round-2 binary)

`tools/r2_configs.sh` (raw JSON lines: gpurun_out/r2_configs.log).  Policy [128, 128], critic [128, 128, 128] everywhere: the
whole-horizon engines cover two-hidden-layer policies of width <= 128 (wide [512]x3 / [436]x3 policies of configs[4] still run the
per-step path, profiles/r1_configs.md).  Initial states: uniform in 0.9 x the feasible state box (obstacle courses: the states
`reset()` drew with their goals and obstacles).

{t3}
Notes
* **Ray-mask rows** report their set exits as `set_exits_reference_quirk_Q2` without failing the gate (SURVEY Q2: the reference's radial
  maps as shipped, ray_mask.py:108-113, drop the paper's factor `safe_dist`, so the executed action of the reference is not confined to
  the safe set; the path reproduces the shipped arithmetic).
* **NavigateSeeker** 65 536 envs x H = 256 and **ManageHousehold** 8192 envs x H = 550 run whole-horizon on the tcgen05 engines (round 1:
  eager loop at 1.5 M env-steps/s, and the 65 536-environment config could not run at all).
* **BalanceQuadrotor + boundary projection** stays solver-bound (lifted 38-variable program per step, one thread per environment,
  profiles/r1_lifted_solver_ncu.md): with box-uniform states almost every guarded step is infeasible and runs to the iteration cap.
  The warp-cooperative solver VERDICT item 4 asks for was not built this round (DESIGN.md section 8).
* **Multi-GPU** (weak scaling, 4096 envs per GPU, `torchrun --nproc-per-node N bench.py --gpus N --steps 30 --warmup 5`, one box, final
  binary): 261.5 / 488.6 / 963.2 / 1892.3 M env-steps/s at 1 / 2 / 4 / 8 GPUs (1.002 / 1.073 / 1.089 / 1.108 ms per episode), efficiency
  0.934 / 0.921 / 0.905; `replicas_in_sync` (policy and critic after a full SHAC iteration) true at every N (the 4-GPU line predates the
  NVML clock sampler).  The per-episode cost of the policy-gradient all-reduce (72 KB, NCCL, between the episode graph and the optimizer
  launch) is ~0.07-0.1 ms: constant, so its share grew as the episode went from 1.98 to 1.00 ms (round 1: 0.916 at 8 GPUs on a 1.98 ms
  episode).  A 2-GPU device timeline (`torchrun ... tools/episode_breakdown.py --gpu-timeline`) names the limiter: the NCCL kernel itself
  is 13 us with 2 us gaps around it; what costs is the HOST side of `torch.distributed.all_reduce`, which makes the host fall ~0.1 ms
  behind the device at the start of the next episode.  With the collective captured as a node of the episode graph
  (`SGBPO_GRAPH_ALLREDUCE=1`) a 2-GPU run measured 502 M env-steps/s, 1.044 ms per episode instead of 485 M, 1.080 ms, replicas in sync
  -- but the process then hung in its exit path (process-group teardown with the graph alive) until the test's timeout, so it stays
  opt-in this round.  Measurement hygiene: the clock sampler used to spawn `nvidia-smi` every 50 ms on rank 0, which enumerates every
  GPU of the box and takes driver locks; through the all-reduce that stalled every rank's timed episodes (2-GPU runs scattered
  between 1.08 and 1.54 ms per episode).  It now reads the same counters in-process through NVML (two consecutive 2-GPU runs: 1.073 /
  1.086 ms).
""")
    par = (G / "r2_parity.txt").read_text()
    (OUT / "r2_parity.md").write_text(f"""# Round 2: measured parity of the whole-horizon engines against the float64 episode

`python tools/parity_report.py` on a B200 (final round-2 binary): the SAME episode (weights, initial states, replayed draws) through the
fused path in float64 on the FMA tiles (the reference value: pinned to the reference's golden episode and to the eager autograd path by
tests/test_gpu_fused_rollout.py) and in float32 on each engine combination.  Entries are relative L2 errors of the rollout buffers over
the whole horizon, |loss - loss64| / |loss64|, and the worst relative L2 error over the policy's parameter gradients.  These are the
numbers behind the bars of tests/test_gpu_rollout_tc.py (engine <= max(3 x the float32 FMA engine's error, 2e-4); headline
instantiations: obs < 5e-3, loss < 2e-3, gradients < 2e-2 after 64 chained steps of an unstable plant).

{par}
All three tensor-core paths use fp16 operands split hi + lo (three products, fp32 accumulation); their per-step error is the
float32 rounding level (1e-7 .. 5e-7 on the buffers), a few 1e-6 on the summed gradients.  The bench line's own gate
(`gates.oracle_slice`, 64 environments of one safeguarded step, forward + backward, against the CPU oracle): max relative error 1.2e-7,
0 verdict mismatches.  Parity stays UNPINNED at the solver boundary (programs P1-P8): cvxpy / cvxpylayers / diffcp / clarabel import
neither here nor on the GPU image (probed with gpurun this round), so the oracle's solutions are checked by KKT / independent LP
solves, not against the reference's solver outputs.
""")
    print("profiles written;", {k: round(v / 1e6, 1) for k, v in per.items()})


if __name__ == "__main__":
    main()
