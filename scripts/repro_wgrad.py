"""Runs one weight-gradient shape per subprocess on the GPU and reports pass / error (debugging aid for tc_wgrad_ts_kernel)."""
import subprocess
import sys

CASES = [  # M, P, Q, q_valid, relu, bias
    (184320, 128, 64, 53, 1, 1), (184320, 128, 64, 64, 1, 1), (184320, 128, 64, 64, 0, 1), (184320, 128, 64, 64, 1, 0),
    (184320, 128, 128, 128, 1, 1), (40000, 128, 64, 64, 1, 1), (184320, 128, 32, 32, 1, 1), (61440, 128, 64, 53, 1, 1),
]


def one(M, P, Q, qv, relu, bias):
    import torch
    sys.path.insert(0, ".")
    from refil_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(M + P + Q)
    dC = torch.randn(M, P, device=dev, generator=g)
    A = torch.randn(M, Q, device=dev, generator=g)
    if qv != Q:
        A[:, qv:] = 0
    y = torch.randn(M, P, device=dev, generator=g)
    dW = torch.zeros(P, qv, device=dev)
    db = torch.zeros(P, device=dev) if bias else None
    for _ in range(3):
        ops.linear_bwd_weight(dC, A, dW, db, relu_y=y if relu else None)
    torch.cuda.synchronize()
    gm = dC.double() * ((y > 0).double() if relu else 1.0)
    ref = 3 * (gm.t() @ A.double())[:, :qv]
    scale = (gm.abs().t() @ A.abs().double()).max().item() * 3
    err = (dW.double() - ref).abs().max().item() / scale
    print("ok rel err %.2e" % err)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        one(*[int(v) for v in sys.argv[1:]])
    else:
        for c in CASES:
            r = subprocess.run([sys.executable, __file__] + [str(v) for v in c], capture_output=True, text=True)
            last = (r.stdout.strip().splitlines() or r.stderr.strip().splitlines() or ["?"])[-1]
            print(c, "->", last[:160], flush=True)
