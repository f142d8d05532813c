"""Measured parity of the whole-horizon engines against the float64 episode (the numbers behind the tolerances of
tests/test_gpu_rollout_tc.py): python tools/parity_report.py > gpurun_out/r2_parity.txt   (needs a GPU)"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import torch  # noqa: E402


def main():
    import test_gpu_rollout_tc as T
    dev = torch.device("cuda", 0)
    torch.set_default_device(dev)
    names = {0: "fma fwd + sequential reverse", 1: "tcgen05 fwd + parallel reverse", 2: "fma fwd + parallel reverse",
             3: "warp-mma fwd + parallel reverse"}
    cases = [("BalanceCartPole", "bp", {"gate": "reference"}, 1024, 64, 240, 0.3),
             ("BalanceCartPole", "bp", {"gate": "always"}, 256, 12, 1000, 0.6),
             ("BalancePendulum", "rm", {"gate": "reference", "zono": True, "linear": True}, 64, 12, 1000, 0.1),
             ("SwingUpCartPole", "rm", {"gate": "always", "zono": False, "linear": False}, 130, 12, 1000, 0.6)]
    print("| case | engine | obs | actions | rewards | loss | worst policy gradient |")
    print("|---|---|---|---|---|---|---|")
    for env_name, sg, opts, N, H, num_steps, spread in cases:
        ref, state, weights, draws = T._episode(env_name, sg, opts, N, H, num_steps, torch.float64, 0, dev, None, None, None, spread)
        for engine in (0, 1, 2, 3):
            got, *_ = T._episode(env_name, sg, opts, N, H, num_steps, torch.float32, engine, dev, state, weights, draws, spread)
            fin = torch.isfinite(ref["obs"]).all(dim=2).all(dim=0) & torch.isfinite(got["obs"]).all(dim=2).all(dim=0)
            e = {k: T._rel(got[k][:, fin], ref[k][:, fin]) for k in ("obs", "actions", "rewards")}
            worst = max(T._rel(got["grads"][k], g) for k, g in ref["grads"].items())
            print(f"| {env_name}+{sg} N={N} H={H} | {names[engine]} | {e['obs']:.1e} | {e['actions']:.1e} | {e['rewards']:.1e} | "
                  f"{abs(got['loss'] - ref['loss']) / abs(ref['loss']):.1e} | {worst:.1e} |")


if __name__ == "__main__":
    main()
