"""GPU parity of the individual sm_100a kernels (through the C ABI) against the CPU oracle / plain torch fp32.
Tolerance: 1e-4 relative (BASELINE.json north_star) -- in practice the fp32-FMA kernels land near 1e-6."""
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _close(a, b, rtol=1e-4, atol=1e-5, what=""):
    a, b = a.detach().cpu().float(), b.detach().cpu().float()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = (a - b).abs().max().item() if a.numel() else 0.0
    assert torch.allclose(a, b, rtol=rtol, atol=atol), (what, err, b.abs().max().item())


@pytest.mark.parametrize("M,N,K,relu", [(1000, 384, 128, False), (777, 64, 128, True), (513, 14, 64, False),
                                        (96, 32, 32, True), (5, 3, 7, False), (2000, 128, 53, True)])
def test_linear_fwd_bwd(M, N, K, relu):
    from refil_b200 import ops
    g = torch.Generator().manual_seed(M + N)
    A, W, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * 0.2, torch.randn(N, generator=g)
    dC = torch.randn(M, N, generator=g)
    Ar, Wr, br = A.clone().requires_grad_(), W.clone().requires_grad_(), b.clone().requires_grad_()
    y = Ar @ Wr.t() + br
    if relu:
        y = torch.relu(y)
    y.backward(dC)
    Ad, Wd, bd, dCd = A.to(DEV), W.to(DEV), b.to(DEV), dC.to(DEV)
    # absolute tolerance relative to the magnitude of the accumulated products (fp32 FFMA ~1e-6, 3xTF32 ~4e-6 of it)
    s_fwd = float((A.abs() @ W.abs().t()).max())
    s_da = float((dC.abs() @ W.abs()).max())
    s_dw = float((dC.abs().t() @ A.abs()).max())
    out = torch.empty(M, N, device=DEV)
    ops.linear_fwd(Ad, Wd, bd, out, relu=relu)
    _close(out, y, atol=5e-6 * s_fwd, what="fwd")
    dA = torch.empty(M, K, device=DEV)
    ops.linear_bwd_data(dCd, Wd, dA, relu_y=out if relu else None)
    _close(dA, Ar.grad, atol=5e-6 * s_da, what="dA")
    dW, db = torch.zeros(N, K, device=DEV), torch.zeros(N, device=DEV)
    ops.linear_bwd_weight(dCd, Ad, dW, db, relu_y=out if relu else None)
    _close(dW, Wr.grad, rtol=2e-4, atol=5e-6 * s_dw, what="dW")
    _close(db, br.grad, rtol=2e-4, atol=2e-4, what="db")


def test_linear_row_mask_and_embed():
    from refil_b200 import ops
    g = torch.Generator().manual_seed(3)
    C, N, na, ne, d, ed, A = 3, 40, 3, 5, 32, 6, 4
    em = (torch.rand(N, ne, generator=g) < 0.3).to(torch.uint8)
    X = torch.randn(C * N * na, d, generator=g)
    W, b = torch.randn(16, d, generator=g) * 0.3, torch.randn(16, generator=g)
    ref = (X @ W.t() + b).view(C, N, na, 16).masked_fill(em[:, :na].bool().view(1, N, na, 1), 0.0).view(-1, 16)
    out = torch.empty(C * N * na, 16, device=DEV)
    ops.linear_fwd(X.to(DEV), W.to(DEV), b.to(DEV), out, row_mask=(em.to(DEV), na, N * na))
    _close(out, ref, what="rowmask")
    # embed over the virtual concat [ents | onehot(la)]
    ents = torch.rand(N, ne, ed, generator=g)
    la = torch.randint(-1, A, (N, ne), generator=g).to(torch.int32)
    W1, b1 = torch.randn(d, ed + A, generator=g) * 0.3, torch.randn(d, generator=g)
    oh = torch.zeros(N, ne, A)
    for k in range(A):
        oh[..., k] = (la == k).float()
    cat = torch.cat([ents, oh], -1).view(N * ne, ed + A)
    W1r, b1r = W1.clone().requires_grad_(), b1.clone().requires_grad_()
    y = torch.relu(cat @ W1r.t() + b1r)
    dY = torch.randn(N * ne, d, generator=g)
    y.backward(dY)
    x1 = torch.empty(N * ne, d, device=DEV)
    ops.embed_fwd(ents.to(DEV), la.to(DEV), A, W1.to(DEV), b1.to(DEV), x1)
    _close(x1, y, what="embed")
    dW, db = torch.zeros(d, ed + A, device=DEV), torch.zeros(d, device=DEV)
    ops.embed_bwd_weight(dY.to(DEV), x1, ents.to(DEV), la.to(DEV), A, dW, db)
    _close(dW, W1r.grad, rtol=2e-4, atol=2e-4, what="embed dW")
    _close(db, b1r.grad, rtol=2e-4, atol=2e-4, what="embed db")


@pytest.mark.parametrize("ne,na,d,H", [(5, 3, 32, 2), (4, 4, 32, 4), (24, 8, 128, 4), (32, 32, 64, 4), (1, 1, 32, 2),
                                      (12, 12, 32, 2), (6, 4, 64, 8), (4, 2, 32, 1), (17, 9, 64, 4),
                                      # 4 heads of 32 with <= 8 query rows: the two-agents-per-lane kernels (attn_*_h4_kernel) with fewer
                                      # than 8 agents, an odd agent count, a single agent, entity counts off the 8-row padding
                                      (16, 5, 128, 4), (9, 8, 128, 4), (8, 3, 128, 4), (24, 1, 128, 4), (31, 6, 128, 4)])
def test_masked_attention_fwd_bwd(ne, na, d, H):
    from oracle import learner_oracle as lo
    from refil_b200 import ops
    g = torch.Generator().manual_seed(ne * 100 + na)
    B, T = 3, 4
    N = B * T
    x = torch.randn(N, ne, d, generator=g)
    w_in = torch.randn(3 * d, d, generator=g) / d ** 0.5
    em = torch.zeros(B, T, ne, dtype=torch.uint8)
    if ne > 2:
        em[0, :, ne - 1] = 1
        em[1, :, na - 1] = 1
    obs = (torch.rand(N, ne, ne, generator=g) < 0.3).to(torch.uint8)
    obs[2] = 1                                        # a unit whose rows are fully masked -> zeros, not NaN
    gbits = (torch.rand(B, ne, generator=g) < 0.5).to(torch.uint8)
    m = lo.imagine_masks(gbits, em[:, 0], na)
    rep = lambda t: t.unsqueeze(1).expand(B, T, na, ne).reshape(N, na, ne)
    emN = em.view(N, ne)
    dflt = (emN[:, :na].bool().unsqueeze(2) | emN.bool().unsqueeze(1)).to(torch.uint8)
    masks = [obs[:, :na], ((obs[:, :na] + rep(m["within"])) > 0).to(torch.uint8),
             ((obs[:, :na] + rep(m["interact"])) > 0).to(torch.uint8)]
    masks2 = [dflt, rep(m["within_noobs"]), rep(m["interact_noobs"])]
    post = torch.zeros(N, na, dtype=torch.uint8)
    eye_w, zero_b = torch.eye(d), torch.zeros(d)
    qkv = (x @ w_in.t()).detach().requires_grad_()          # leaf: the kernel's input is QKV itself
    eye3 = torch.eye(3 * d)
    douts = torch.randn(3, N, na, d, generator=g)
    for name, mlist, copies in (
            ("agent", masks, [(obs.to(DEV), ne * ne, 0), (obs.to(DEV), ne * ne, 1), (obs.to(DEV), ne * ne, 2)]),
            ("mixer", masks2, [(None, 0, 8), (None, 0, 1 | 4), (None, 0, 2 | 4)])):
        refs = [lo.entity_attention(qkv, eye3, eye_w, zero_b, mk, post, H) for mk in mlist]
        out = torch.empty(3, N, na, d, device=DEV)
        qkv_d = qkv.detach().reshape(N * ne, 3 * d).to(DEV)
        ops.masked_attn_fwd(qkv_d, out, copies, gbits.to(DEV), emN.to(DEV), N, T, ne, na, d, H)
        for c in range(3):
            _close(out[c], refs[c], what="%s fwd copy %d" % (name, c))
        assert torch.isfinite(out).all()
        loss = sum((r * douts[c]).sum() for c, r in enumerate(refs))
        (dqkv_ref,) = torch.autograd.grad(loss, qkv)
        dqkv = torch.empty(N * ne, 3 * d, device=DEV)
        ops.masked_attn_bwd(qkv_d, douts.to(DEV), dqkv, copies, gbits.to(DEV), emN.to(DEV), N, T, ne, na, d, H)
        _close(dqkv, dqkv_ref.reshape(N * ne, 3 * d), what="%s dqkv" % name)


@pytest.mark.parametrize("r,na,CB,T", [(16, 3, 9, 6), (64, 8, 5, 7), (32, 1, 1, 1), (128, 2, 3, 3),
                                       (64, 8, 48, 12),      # 384 sequences = the 16-episode shard: 24 CTAs of the mma scan
                                       (64, 8, 300, 5),      # 2400 sequences: two m16 tiles per CTA
                                       (32, 5, 7, 9), (64, 3, 7, 20)])   # ragged: 35 / 21 sequences (partial m16 tiles)
def test_gru_scan_fwd_bwd(r, na, CB, T):
    from oracle import learner_oracle as lo
    from refil_b200 import ops
    g = torch.Generator().manual_seed(r + na)
    R = CB * T * na
    gi = torch.randn(R, 3 * r, generator=g)
    whh = (torch.randn(3 * r, r, generator=g) / r ** 0.5)
    bhh = torch.randn(3 * r, generator=g) * 0.1
    h0 = torch.randn(CB * na, r, generator=g) * 0.5
    dhs = torch.randn(R, r, generator=g)
    gir, whr, bhr = gi.clone().requires_grad_(), whh.clone().requires_grad_(), bhh.clone().requires_grad_()
    gi4 = gir.view(CB, T, na, 3 * r)
    h = h0.clone()
    hs = []
    wih_eye = torch.zeros(3 * r, 3 * r)          # feed gi through gru_cell: x = gi, W_ih = I, b_ih = 0
    for t in range(T):
        x = gi4[:, t].reshape(CB * na, 3 * r)
        gh = h @ whr.t() + bhr
        i_r, i_z, i_n = x.chunk(3, 1)
        h_r, h_z, h_n = gh.chunk(3, 1)
        rg, zg = torch.sigmoid(i_r + h_r), torch.sigmoid(i_z + h_z)
        ng = torch.tanh(i_n + rg * h_n)
        h = (1 - zg) * ng + zg * h
        hs.append(h.view(CB, na, r))
    hs_ref = torch.stack(hs, 1).reshape(R, r)
    (hs_ref * dhs).sum().backward()
    HS, gates = torch.empty(R, r, device=DEV), torch.empty(R, 4 * r, device=DEV)
    ops.gru_scan_fwd(gi.to(DEV), whh.to(DEV), bhh.to(DEV), h0.to(DEV), HS, gates, CB * na, T, na)
    _close(HS, hs_ref, what="hs")
    dGI, dGH = torch.empty(R, 3 * r, device=DEV), torch.empty(R, 3 * r, device=DEV)
    ops.gru_scan_bwd(dhs.to(DEV), gates, HS, h0.to(DEV), whh.to(DEV), dGI, dGH, CB * na, T, na)
    _close(dGI, gir.grad, what="dgi")
    # recurrent weight gradient from dGH and the shifted state stack (h0 = 0 in training; use zeros here)
    HS0, gates0 = torch.empty(R, r, device=DEV), torch.empty(R, 4 * r, device=DEV)
    ops.gru_scan_fwd(gi.to(DEV), whh.to(DEV), bhh.to(DEV), None, HS0, gates0, CB * na, T, na)
    ops.gru_scan_bwd(dhs.to(DEV), gates0, HS0, None, whh.to(DEV), dGI, dGH, CB * na, T, na)
    dW, db = torch.zeros(3 * r, r, device=DEV), torch.zeros(3 * r, device=DEV)
    ops.gru_bwd_weight_hh(dGH, HS0, na, T, dW, db)
    gir2, whr2, bhr2 = gi.clone().requires_grad_(), whh.clone().requires_grad_(), bhh.clone().requires_grad_()
    h = torch.zeros(CB * na, r)
    hs = []
    for t in range(T):
        h = lo.gru_cell(gir2.view(CB, T, na, 3 * r)[:, t].reshape(CB * na, 3 * r), h, torch.eye(3 * r), whr2,
                        torch.zeros(3 * r), bhr2)
        hs.append(h.view(CB, na, r))
    (torch.stack(hs, 1).reshape(R, r) * dhs).sum().backward()
    _close(dW, whr2.grad, rtol=2e-4, atol=2e-4, what="dWhh")
    _close(db, bhr2.grad, rtol=2e-4, atol=2e-4, what="dbhh")


@pytest.mark.parametrize("kind,imagine,softmax,tanh", [("flex_qmix", True, True, False), ("flex_qmix", False, True, False),
                                                       ("flex_qmix", True, False, True), ("lin_flex_qmix", True, True, False),
                                                       ("lin_flex_qmix", False, False, False), ("vdn", True, True, False)])
def test_mixer_combine_fwd_bwd(kind, imagine, softmax, tanh):
    import torch.nn.functional as F
    from refil_b200 import ops
    g = torch.Generator().manual_seed(11)
    N, na, me = 37, 3, 8
    Cw = 3 if imagine else 1
    W1 = torch.randn(Cw, N, na, me, generator=g).requires_grad_()
    B1, WF, V = [torch.randn(N, na, me, generator=g).requires_grad_() for _ in range(3)]
    q, qW, qI = [torch.randn(N, na, generator=g).requires_grad_() for _ in range(3)]
    gp, gi = torch.randn(N, generator=g), torch.randn(N, generator=g)

    def mixw(x, dim):
        return torch.softmax(x, dim) if softmax else x.abs()

    def flex(qs, w1raw):
        w1 = mixw(w1raw, -1)
        pre = torch.bmm(qs.unsqueeze(1), w1) + B1.mean(1).unsqueeze(1)
        hid = torch.tanh(pre) if tanh else F.elu(pre)
        wf = mixw(WF.mean(1), -1)
        return (torch.bmm(hid, wf.unsqueeze(2)).view(N) + V.mean((1, 2)))

    def lin(qs, w1raw):
        w1 = mixw(w1raw.mean(-1), 1)
        return (qs * w1).sum(1) + V.mean((1, 2))

    if kind == "vdn":
        y, y_im = q.sum(1), torch.cat([qW, qI], 1).sum(1)
    else:
        f = flex if kind == "flex_qmix" else lin
        y = f(q, W1[0])
        y_im = f(torch.cat([qW, qI], 1), torch.cat([W1[1], W1[2]], 1)) if imagine else None
    loss = (y * gp).sum() + ((y_im * gi).sum() if imagine else 0.0)
    loss.backward()
    d = lambda t: t.detach().to(DEV).contiguous()
    qtot, qtot_im = torch.empty(N, device=DEV), torch.empty(N, device=DEV)
    k = ops.MIX_KIND[kind]
    ops.mixer_fwd(k, d(W1), d(B1), d(WF), d(V), d(q), d(qW), d(qI), qtot, qtot_im if imagine else None, N, na, me, Cw,
                  imagine, softmax, tanh)
    _close(qtot, y, what="qtot")
    if imagine:
        _close(qtot_im, y_im, what="qtot_im")
    dW1, dB1, dWF, dV = [torch.full_like(d(t), 7.0) for t in (W1, B1, WF, V)]
    dq = torch.full((3, N, na), 7.0, device=DEV)
    ops.mixer_bwd(k, d(W1), d(B1), d(WF), d(V), d(q), d(qW), d(qI), d(gp), d(gi) if imagine else None, dW1, dB1, dWF, dV,
                  dq[0], dq[1] if imagine else None, dq[2] if imagine else None, N, na, me, Cw, imagine, softmax, tanh)
    _close(dq[0], q.grad, what="dq")
    if imagine:
        _close(dq[1], qW.grad, what="dqW")
        _close(dq[2], qI.grad, what="dqI")
    if kind != "vdn":
        _close(dW1, W1.grad, what="dW1")
        _close(dV, V.grad, what="dV")
    if kind == "flex_qmix":
        _close(dB1, B1.grad, what="dB1")
        _close(dWF, WF.grad, what="dWF")


def test_select_actions_first_max_and_epsilon():
    from refil_b200 import ops
    g = torch.Generator().manual_seed(5)
    B, na, A = 64, 4, 6
    q = torch.randint(0, 3, (B, na, A), generator=g).float()          # many ties -> first-max rule matters
    avail = (torch.rand(B, na, A, generator=g) < 0.6).to(torch.int32)
    avail[..., 1] = 1
    mq = q.clone()
    mq[avail == 0] = -float("inf")
    ref = mq.max(dim=2)[1]
    out = torch.zeros(B, na, dtype=torch.int64, device=DEV)
    ops.select_actions(q.to(DEV), avail.to(DEV), None, None, None, 0.0, out, B, na, A)
    assert torch.equal(out.cpu(), ref)
    u = torch.rand(2, B, na, device=DEV)
    ops.select_actions(q.to(DEV), avail.to(DEV), u[0].contiguous(), u[1].contiguous(), None, 1.0, out, B, na, A)
    picked = torch.gather(avail, 2, out.cpu().unsqueeze(-1))
    assert (picked == 1).all()                                         # exploration only picks available actions


def test_missing_cuda_fails_loudly():
    from refil_b200 import _lib, ops
    with pytest.raises(_lib.RefilError):
        ops.linear_fwd(torch.zeros(4, 4), torch.zeros(4, 4), None, torch.zeros(4, 4))


@pytest.mark.parametrize("M,N,K", [(1000, 384, 128), (300, 64, 64), (4096, 192, 64), (129, 128, 128), (20000, 32, 128),
                                   (513, 256, 32), (2048, 32, 64)])
def test_tc_gemm_3xtf32_forward(M, N, K):
    """tcgen05 3xTF32 GEMM vs float64: fp32-grade accuracy (a single TF32 pass would be ~1e-3)."""
    from refil_b200 import _lib, ops
    assert _lib.load().refil_tc_gemm_supported(M, N, K)
    g = torch.Generator().manual_seed(M + N + K)
    A, W, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * 0.3, torch.randn(N, generator=g)
    C, na, ne = 1, 4, 6
    em = None
    if M % na == 0:
        em = (torch.rand(M // na, ne, generator=g) < 0.3).to(torch.uint8)
    ref = A.double() @ W.double().t() + b.double()
    ref = torch.relu(ref)
    if em is not None:
        ref = ref.view(M // na, na, N).masked_fill(em[:, :na].bool().unsqueeze(-1), 0.0).view(M, N)
    out = torch.full((M, N), float("nan"), device=DEV)
    ops.tc_gemm_tn(A.to(DEV), W.to(DEV), K, 1, out, N, K, bias=b.to(DEV), relu=True,
                   c_row_mask=(em.to(DEV), na, M) if em is not None else None)
    torch.cuda.synchronize()
    err = (out.cpu().double() - ref).abs().max().item()
    scale = (A.abs().double() @ W.abs().double().t()).max().item()
    assert err <= 4e-6 * scale, (err, scale)


@pytest.mark.parametrize("M,N,K", [(1000, 384, 128), (777 * 4, 64, 128), (2048, 128, 128), (600, 192, 64)])
def test_tc_gemm_3xtf32_backward_data(M, N, K):
    """dA = (dC * relu'(y) * rowmask) W through the strided-B form, vs float64."""
    from refil_b200 import ops
    g = torch.Generator().manual_seed(M + N)
    dC, W = torch.randn(M, N, generator=g), torch.randn(N, K, generator=g) * 0.3
    y = torch.randn(M, N, generator=g)
    na, ne = 4, 5
    em = (torch.rand(M // na, ne, generator=g) < 0.3).to(torch.uint8)
    gmat = dC.double() * (y > 0).double()
    gmat = gmat.view(M // na, na, N).masked_fill(em[:, :na].bool().unsqueeze(-1), 0.0).view(M, N)
    ref = gmat @ W.double()
    dA = torch.full((M, K), float("nan"), device=DEV)
    ops.linear_bwd_data(dC.to(DEV), W.to(DEV), dA, relu_y=y.to(DEV), row_mask=(em.to(DEV), na, M))
    torch.cuda.synchronize()
    err = (dA.cpu().double() - ref).abs().max().item()
    scale = (gmat.abs() @ W.abs().double()).max().item()
    assert err <= 4e-6 * scale, (err, scale)


@pytest.mark.parametrize("M,P,Q", [(5000, 384, 128), (777 * 4, 64, 128), (4096, 128, 128), (600, 192, 64), (40000, 32, 128),
                                   (1000, 128, 32), (333 * 4, 20, 64)])
def test_tc_wgrad_3xtf32(M, P, Q):
    """dW = (dC * relu'(y) * rowmask)^T A and db = colsum, on tcgen05 with MN-major operands, vs float64."""
    from refil_b200 import _lib, ops
    assert _lib.load().refil_tc_wgrad_supported(M, P, Q)
    g = torch.Generator().manual_seed(M + P)
    dC, A = torch.randn(M, P, generator=g), torch.randn(M, Q, generator=g)
    y = torch.randn(M, P, generator=g)
    na, ne = 4, 5
    em = (torch.rand(M // na, ne, generator=g) < 0.3).to(torch.uint8)
    gmat = dC.double() * (y > 0).double()
    gmat = gmat.view(M // na, na, P).masked_fill(em[:, :na].bool().unsqueeze(-1), 0.0).view(M, P)
    ref_w, ref_b = gmat.t() @ A.double(), gmat.sum(0)
    dW, db = torch.ones(P, Q, device=DEV), torch.ones(P, device=DEV)          # accumulates into existing values
    ops.linear_bwd_weight(dC.to(DEV), A.to(DEV), dW, db, relu_y=y.to(DEV), row_mask=(em.to(DEV), na, M))
    torch.cuda.synchronize()
    scale = (gmat.abs().t() @ A.abs().double()).max().item()
    err = (dW.cpu().double() - 1.0 - ref_w).abs().max().item()
    assert err <= 4e-6 * scale, ("dW", err, scale)
    errb = (db.cpu().double() - 1.0 - ref_b).abs().max().item()
    assert errb <= 4e-6 * gmat.abs().sum(0).max().item(), ("db", errb)


@pytest.mark.parametrize("M,P,Q", [(5000, 384, 128), (3000, 64, 32)])
def test_tc_wgrad_without_bias_gradient(M, P, Q):
    """db = None (in_trans has no bias): the Y tile carries no ones-atom, N = Q exactly."""
    from refil_b200 import ops
    g = torch.Generator().manual_seed(M + Q)
    dC, A = torch.randn(M, P, generator=g), torch.randn(M, Q, generator=g)
    ref_w = dC.double().t() @ A.double()
    dW = torch.zeros(P, Q, device=DEV)
    ops.linear_bwd_weight(dC.to(DEV), A.to(DEV), dW, None)
    torch.cuda.synchronize()
    scale = (dC.abs().double().t() @ A.abs().double()).max().item()
    err = (dW.cpu().double() - ref_w).abs().max().item()
    assert err <= 4e-6 * scale, (err, scale)


# ---- full north-star sizes (B=128, T=60, ne=24, d=128): size-independent properties + float64 spot checks ------------------
NS_ROWS = 128 * 60 * 24


def test_full_size_dense_forward_against_float64_rows():
    """in_trans at the benchmark size on the tensor-memory path: 4096 sampled rows against float64, and the whole output against
    the independent fp32 FFMA kernel."""
    from refil_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(1)
    M, N, K = NS_ROWS, 384, 128
    A = torch.randn(M, K, device=DEV, generator=g)
    W = torch.randn(N, K, device=DEV, generator=g) * 0.2
    out_tc = torch.empty(M, N, device=DEV)
    ops.linear_fwd(A, W, None, out_tc)
    rows = torch.randint(0, M, (4096,), device=DEV, generator=g)
    ref = A[rows].double() @ W.double().t()
    scale = (A[rows].abs().double() @ W.abs().double().t()).max().item()
    assert (out_tc[rows].double() - ref).abs().max().item() <= 4e-6 * scale
    old = ops.USE_TENSOR_CORES
    try:
        ops.USE_TENSOR_CORES = False
        out_ff = torch.empty(M, N, device=DEV)
        ops.linear_fwd(A, W, None, out_ff)
    finally:
        ops.USE_TENSOR_CORES = old
    assert (out_tc - out_ff).abs().max().item() <= 1e-5 * scale


def test_full_size_split_k_backward_and_weight_gradient():
    """Backward-data of in_trans (K = 3d, k-sliced reduce-add) is linear in the upstream gradient, and the weight gradient at
    full size matches float64."""
    from refil_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(2)
    M, N, K = NS_ROWS, 384, 128
    W = torch.randn(N, K, device=DEV, generator=g) * 0.2
    d1 = torch.randn(M, N, device=DEV, generator=g)
    d2 = torch.randn(M, N, device=DEV, generator=g)
    o1, o2, o12 = (torch.empty(M, K, device=DEV) for _ in range(3))
    ops.linear_bwd_data(d1, W, o1)
    ops.linear_bwd_data(d2, W, o2)
    ops.linear_bwd_data(d1 + d2, W, o12)
    scale = o12.abs().max().item()
    assert (o12 - (o1 + o2)).abs().max().item() <= 2e-5 * scale
    X = torch.randn(M, K, device=DEV, generator=g)
    dW = torch.zeros(N, K, device=DEV)
    ops.linear_bwd_weight(d1, X, dW, None)
    ref = d1.double().t() @ X.double()
    scale = (d1.abs().double().t() @ X.abs().double()).max().item()      # fp32 accumulation over 184 320 rows
    assert (dW.double() - ref).abs().max().item() <= 4e-6 * scale


def test_full_size_attention_copies_are_independent():
    """One launch with three mask copies equals three single-copy launches (forward and backward), at the benchmark size."""
    from refil_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(3)
    B, T, ne, na, d, H = 128, 60, 24, 8, 128, 4
    N = B * T
    qkv = torch.randn(N * ne, 3 * d, device=DEV, generator=g)
    obs = (torch.rand(N, ne, ne, device=DEV, generator=g) < 0.2).to(torch.uint8)
    em = (torch.rand(B, 1, ne, device=DEV, generator=g) < 0.1).to(torch.uint8).expand(B, T, ne).contiguous().view(N, ne)
    gb = (torch.rand(B, ne, device=DEV, generator=g) < 0.5).to(torch.uint8)
    copies = [(obs, ne * ne, 0), (obs, ne * ne, 1), (obs, ne * ne, 2)]
    out3 = torch.empty(3, N, na, d, device=DEV)
    ops.masked_attn_fwd(qkv, out3, copies, gb, em, N, T, ne, na, d, H)
    dout = torch.randn(3, N, na, d, device=DEV, generator=g)
    dq3 = torch.empty(N * ne, 3 * d, device=DEV)
    ops.masked_attn_bwd(qkv, dout, dq3, copies, gb, em, N, T, ne, na, d, H)
    acc = torch.zeros_like(dq3)
    for c in range(3):
        o1 = torch.empty(1, N, na, d, device=DEV)
        ops.masked_attn_fwd(qkv, o1, [copies[c]], gb, em, N, T, ne, na, d, H)
        assert torch.equal(o1[0], out3[c])
        dq1 = torch.empty(N * ne, 3 * d, device=DEV)
        ops.masked_attn_bwd(qkv, dout[c:c + 1].contiguous(), dq1, [copies[c]], gb, em, N, T, ne, na, d, H)
        acc += dq1
    assert torch.isfinite(out3).all()
    assert (acc - dq3).abs().max().item() <= 1e-4 * dq3.abs().max().item()


def test_grouped_dense_launches_match_single_launches():
    """Grouped tensor-core launches (one kernel, blockIdx.y = problem; include/refil_b200.h RefilGemmDesc / RefilWgradDesc): the
    same layer of several networks with different operands and ROW COUNTS, against fp64 references -- forward with bias / relu /
    row mask, backward-data (W^T by strides, k-sliced reduction for K = 3d), weight + bias gradients with a fused relu' mask,
    and the in-place zero-padded operands (fc1: 53-column weight against a 64-column packed input)."""
    from refil_b200 import ops
    g = torch.Generator().manual_seed(77)
    l0 = ops.launch_count()
    Ms = [640, 1000, 384, 2048, 300]
    # forward: N = 128, K = 64 with a 53-column weight (zero-padded reduction), relu
    K, Kw, N = 64, 53, 128
    As = [torch.randn(M, K, generator=g) for M in Ms]
    for A in As:
        A[:, Kw:] = 0
    Ws = [torch.randn(N, Kw, generator=g) * 0.2 for _ in Ms]
    bs = [torch.randn(N, generator=g) for _ in Ms]
    outs = [torch.empty(M, N, device=DEV) for M in Ms]
    ops.linear_fwd_group([(A.to(DEV), W.to(DEV), b.to(DEV), o, True, None) for A, W, b, o in zip(As, Ws, bs, outs)])
    assert ops.launch_count() - l0 == 1
    for A, W, b, o in zip(As, Ws, bs, outs):
        ref = torch.relu(A[:, :Kw].double() @ W.double().t() + b.double()).float()
        _close(o, ref, atol=5e-6 * float((A.abs() @ torch.nn.functional.pad(W, (0, K - Kw)).abs().t()).max()), what="group fwd")
    # weight gradients into the 53-column weight, relu' fused
    dCs = [torch.randn(M, N, generator=g) for M in Ms]
    dWs = [torch.zeros(N, Kw, device=DEV) for _ in Ms]
    dbs = [torch.zeros(N, device=DEV) for _ in Ms]
    l1 = ops.launch_count()
    ops.linear_bwd_weight_group([(dC.to(DEV), A.to(DEV), dW, db, o, None) for dC, A, dW, db, o in zip(dCs, As, dWs, dbs, outs)])
    assert ops.launch_count() - l1 == 1
    for dC, A, dW, db, o in zip(dCs, As, dWs, dbs, outs):
        gm = dC.double() * (o.cpu().double() > 0)
        _close(dW, (gm.t() @ A[:, :Kw].double()).float(), rtol=2e-4, atol=5e-6 * float((dC.abs().t() @ A.abs()).max()), what="group dW")
        _close(db, gm.sum(0).float(), rtol=2e-4, atol=2e-4, what="group db")
    # backward-data with a k-sliced reduction (N = 384 -> K = 128, the in_trans shape): dA = dC W
    N2, K2 = 384, 128
    dC2 = [torch.randn(M, N2, generator=g) for M in Ms]
    W2 = [torch.randn(N2, K2, generator=g) * 0.1 for _ in Ms]
    dA2 = [torch.empty(M, K2, device=DEV) for M in Ms]
    l2 = ops.launch_count()
    ops.linear_bwd_data_group([(dC.to(DEV), W.to(DEV), dA, None, None) for dC, W, dA in zip(dC2, W2, dA2)])
    assert ops.launch_count() - l2 == 1
    for dC, W, dA in zip(dC2, W2, dA2):
        _close(dA, (dC.double() @ W.double()).float(), atol=5e-6 * float((dC.abs() @ W.abs()).max()), what="group dA")
    # row masks on the output side, 9 problems -> two launches (8 + 1)
    na, ne, Nn, C = 3, 5, 96, 1
    em = (torch.rand(Nn, ne, generator=g) < 0.3).to(torch.uint8)
    X = [torch.randn(C * Nn * na, 128, generator=g) for _ in range(9)]
    W3 = [torch.randn(32, 128, generator=g) * 0.3 for _ in range(9)]
    b3 = [torch.randn(32, generator=g) for _ in range(9)]
    o3 = [torch.empty(C * Nn * na, 32, device=DEV) for _ in range(9)]
    l3 = ops.launch_count()
    rm = (em.to(DEV), na, Nn * na)
    ops.linear_fwd_group([(x.to(DEV), w.to(DEV), b.to(DEV), o, False, rm) for x, w, b, o in zip(X, W3, b3, o3)])
    assert ops.launch_count() - l3 == 2
    for x, w, b, o in zip(X, W3, b3, o3):
        ref = (x @ w.t() + b).view(C, Nn, na, 32).masked_fill(em[:, :na].bool().view(1, Nn, na, 1), 0.0).view(-1, 32)
        _close(o, ref, atol=1e-4, what="group rowmask")


@pytest.mark.gpu
@pytest.mark.parametrize("units,ne,na", [(200, 24, 8), (61, 16, 4)])
def test_row_grouped_dense_product_touches_only_the_agent_rows(units, ne, na):
    """refil_tc_gemm_tn_rows: C[g*ne + i] (+)= A[g*ne + i] W^T for i < na only -- the Q third of in_trans on the agent rows
    (attention.py:46-48).  The other rows of C keep their (NaN) contents; accumulate adds into C."""
    from refil_b200 import ops
    d = 128
    g = torch.Generator().manual_seed(units)
    A = torch.randn(units * ne, d, generator=g)
    W = torch.randn(d, d, generator=g) * 0.2
    C = torch.full((units * ne, 3 * d), float("nan"), device=DEV)
    ops.tc_gemm_tn_rows(A.to(DEV), W.to(DEV), C, units, na, ne)
    torch.cuda.synchronize()
    ref = (A.double() @ W.double().t()).view(units, ne, d)
    got = C.cpu().view(units, ne, 3 * d)
    scale = (A.abs().double() @ W.abs().double().t()).max().item()
    assert (got[:, :na, :d].double() - ref[:, :na]).abs().max().item() <= 4e-6 * scale
    assert torch.isnan(got[:, na:]).all() and torch.isnan(got[:, :na, d:]).all()
    # accumulate, transposed weight (the backward-data form) into a column slice
    D = torch.ones(units * ne, d, device=DEV)
    G = torch.randn(units * ne, 3 * d, generator=g)
    ops.tc_gemm_tn_rows(G.to(DEV)[:, :d], W.to(DEV), D, units, na, ne, accumulate=True, transpose_w=True)
    torch.cuda.synchronize()
    refd = (G[:, :d].double() @ W.double()).view(units, ne, d)
    gd = D.cpu().view(units, ne, d)
    scale = (G[:, :d].abs().double() @ W.abs().double()).max().item()
    assert (gd[:, :na].double() - 1.0 - refd[:, :na]).abs().max().item() <= 4e-6 * scale
    assert (gd[:, na:] == 1.0).all()


@pytest.mark.gpu
def test_split_in_trans_feeds_the_attention_kernels_like_the_full_product():
    """K|V for all rows + Q for the agent rows (Q of the other rows left as NaN) gives the same attention output and the same
    backward as the full in_trans product."""
    from refil_b200 import ops
    units, ne, na, d, H, T = 96, 24, 8, 128, 4, 8
    g = torch.Generator().manual_seed(7)
    x1 = torch.randn(units * ne, d, generator=g).to(DEV)
    W = (torch.randn(3 * d, d, generator=g) * 0.2).to(DEV)
    full = torch.empty(units * ne, 3 * d, device=DEV)
    ops.linear_fwd(x1, W, None, full)
    split = torch.full((units * ne, 3 * d), float("nan"), device=DEV)
    ops.in_trans_fwd_split(x1, W, split, units, ne, na)
    em = (torch.rand(units, ne, generator=g) < 0.2).to(torch.uint8).to(DEV)
    om = (torch.rand(units, ne, ne, generator=g) < 0.3).to(torch.uint8).to(DEV)
    copies = [(om, ne * ne, 0)]
    outs, grads = [], []
    dout = torch.randn(1, units, na, d, generator=g).to(DEV)
    for qkv in (full, split):
        out = torch.empty(1 * units * na, d, device=DEV)
        ops.masked_attn_fwd(qkv, out, copies, None, em, units, T, ne, na, d, H)
        dqkv = torch.empty(units * ne, 3 * d, device=DEV)
        ops.masked_attn_bwd(qkv, dout.view(-1, d), dqkv, copies, None, em, units, T, ne, na, d, H)
        outs.append(out)
        grads.append(dqkv)
    torch.cuda.synchronize()
    assert torch.isfinite(outs[1]).all() and torch.isfinite(grads[1]).all()
    assert (outs[0] - outs[1]).abs().max().item() <= 1e-5 * outs[0].abs().max().item()
    assert (grads[0] - grads[1]).abs().max().item() <= 1e-5 * grads[0].abs().max().item()


@pytest.mark.gpu
@pytest.mark.parametrize("P,Q,qv", [(128, 64, 53), (128, 128, 128), (32, 128, 128)])
def test_full_size_weight_gradient_with_relu_mask_and_bias(P, Q, qv):
    """fc1 / fc2-shaped weight gradients at the benchmark's row count with the relu' mask and the bias gradient: 39 row chunks per CTA
    on every SM -- the regime in which an odd-depth raw Y ring let a producer group pass an mbarrier phase early (launch failure at
    full size only; scripts/repro_wgrad.py).  Checked against float64."""
    from refil_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(P + Q)
    M = NS_ROWS
    dC = torch.randn(M, P, device=DEV, generator=g)
    A = torch.randn(M, Q, device=DEV, generator=g)
    if qv != Q:
        A[:, qv:] = 0
    y = torch.randn(M, P, device=DEV, generator=g)
    dW, db = torch.zeros(P, qv, device=DEV), torch.zeros(P, device=DEV)
    for _ in range(2):
        ops.linear_bwd_weight(dC, A, dW, db, relu_y=y)
    torch.cuda.synchronize()
    gm = dC.double() * (y > 0).double()
    ref = 2 * (gm.t() @ A.double())[:, :qv]
    scale = 2 * (gm.abs().t() @ A.abs().double()).max().item()
    assert (dW.double() - ref).abs().max().item() <= 4e-6 * scale
    assert (db.double() - 2 * gm.sum(0)).abs().max().item() <= 4e-6 * 2 * gm.abs().sum(0).max().item()


@pytest.mark.gpu
def test_row_grouped_weight_gradient_matches_float64():
    """in_trans weight gradient split: the K|V rows of dW from every entity row, the Q rows from the agent rows only (rank-3 tensor maps
    for X and Y), one launch of two problems."""
    from refil_b200 import ops
    units, ne, na, d = 300, 24, 8, 128
    g = torch.Generator(device=DEV).manual_seed(11)
    dqkv = torch.randn(units * ne, 3 * d, device=DEV, generator=g)
    dqkv.view(units, ne, 3 * d)[:, na:, :d] = float("nan")        # never read: the split takes the Q columns of the agent rows only
    x1 = torch.randn(units * ne, d, device=DEV, generator=g)
    dW = torch.ones(3 * d, d, device=DEV)
    ops.in_trans_bwd_weight_split(dqkv, x1, dW, units, ne, na)
    torch.cuda.synchronize()
    clean = torch.nan_to_num(dqkv, nan=0.0)
    ref = clean.double().t() @ x1.double()
    scale = (clean.abs().double().t() @ x1.abs().double()).max().item()
    assert torch.isfinite(dW).all()
    assert (dW.double() - 1.0 - ref).abs().max().item() <= 4e-6 * scale
