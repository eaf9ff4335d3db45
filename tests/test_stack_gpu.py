"""EGTStack: the reference's layer loop (graph_xformer_model_base.py:335-341) with every weight of every layer in ONE
flat parameter -- same numbers as the layers run one by one, and one gradient buffer for the all-reduce."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.parametrize('ffn', [False, True])
def test_stack_equals_layers_and_has_one_gradient_buffer(ffn):
    import egt_b200
    torch.manual_seed(0)
    B, N, d, de, nh, L = 3, 40, 64, 8, 8, 3
    kw = dict(model_width=d, edge_width=de, num_heads=nh, scale_degree=True)
    stack = egt_b200.EGTStack(L, ffn=ffn, **kw).to(DEV)
    with torch.no_grad():
        stack.flat.add_(0.02 * torch.randn_like(stack.flat))
    singles = [(egt_b200.EGTLayer(tag=f'{i:0>2d}', **kw) if ffn else egt_b200.EGTBlock(tag=f'{i:0>2d}', **kw)).to(DEV) for i in range(L)]
    # copy the stack's weights into stand-alone layers
    k = 0
    for sl in singles:
        mods = [sl] if not ffn else [sl.block, sl.ffn_node, sl.ffn_edge]
        for m in mods:
            with torch.no_grad():
                m.flat.copy_(stack._owners[k].flat)
            k += 1
    mask = (torch.arange(N)[None] < torch.tensor([40, 31, 7])[:, None]).to(DEV)
    h = torch.randn(B, N, d, device=DEV).bfloat16()
    e = torch.randn(B, N, N, de, device=DEV).bfloat16()
    hs, es = h.clone().requires_grad_(True), e.clone().requires_grad_(True)
    h2, e2 = stack(hs, es, mask)
    hr, er = h.clone().requires_grad_(True), e.clone().requires_grad_(True)
    x, y = hr, er
    for sl in singles:
        x, y = sl(x, y, mask)
    assert torch.equal(h2, x) and torch.equal(e2, y)
    dh, dE = torch.randn_like(h2), torch.randn_like(e2)
    torch.autograd.backward([h2, e2], [dh, dE])
    torch.autograd.backward([x, y], [dh, dE])
    assert torch.equal(hs.grad, hr.grad) and torch.equal(es.grad, er.grad)
    # ONE gradient buffer: every module's gradient is a slice of stack.flat.grad
    assert stack.flat.grad is not None and sum(p.numel() for p in stack.parameters()) == stack.flat.numel()
    k = 0
    for sl in singles:
        mods = [sl] if not ffn else [sl.block, sl.ffn_node, sl.ffn_edge]
        for m in mods:
            off, n = stack._spans[k]
            ref = m.flat.grad
            got = stack.flat.grad[off:off + n]
            torch.testing.assert_close(got, ref, rtol=1e-3, atol=1e-3 * float(ref.abs().max()))   # atomics: summation order
            k += 1


def test_full_layer_step_replays_from_a_cuda_graph():
    """One full layer (attention block + node FFN + edge FFN) forward + backward captured after eager warm-up steps:
    the replay reproduces the eager gradients (bench.py's full_layer leg relies on this)."""
    import egt_b200
    torch.manual_seed(0)
    B, N, d, de = 2, 48, 64, 8
    lay = egt_b200.EGTStack(1, ffn=True, model_width=d, edge_width=de, num_heads=8).to(DEV)
    lay.train(False)
    h = torch.randn(B, N, d, device=DEV).bfloat16()
    e = torch.randn(B, N, N, de, device=DEV).bfloat16()
    m = torch.ones(B, N, dtype=torch.bool, device=DEV)
    dh, de_ = torch.randn_like(h), torch.randn_like(e)

    def step():
        hh, ee = h.detach().requires_grad_(True), e.detach().requires_grad_(True)
        lay.flat.grad = None
        h2, e2 = lay(hh, ee, m)
        torch.autograd.backward([h2, e2], [dh, de_])
        return hh.grad, ee.grad, lay.flat.grad

    for _ in range(2):
        ref = [t.clone() for t in step()]          # eager, on the default stream
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = step()
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    for got, want in zip(out, ref):
        scale = float(want.float().abs().max())
        assert float((got.float() - want.float()).abs().max()) <= 2e-3 * scale   # weight gradients: atomics, order differs


@pytest.mark.parametrize('d,de,nh,N', [(64, 8, 8, 100), (128, 32, 16, 70)])
def test_training_step_with_random_masks_replays_from_a_cuda_graph(d, de, nh, N):
    """random_mask_prob > 0 in a captured step: the kernels add a device counter to their Philox offset and the captured
    step bumps it, so replay r draws the mask of eager call (offset at capture + r) -- bit-exactly, forward and backward --
    and consecutive replays differ (reference: fresh noise every step, egt_layers.py:103-108)."""
    import egt_b200
    torch.manual_seed(1)
    B = 3
    blk = egt_b200.EGTBlock(model_width=d, edge_width=de, num_heads=nh, random_mask_prob=0.3, seed=77).to(DEV)
    blk.train(True)
    h = torch.randn(B, N, d, device=DEV).bfloat16()
    e = torch.randn(B, N, N, de, device=DEV).bfloat16()
    m = torch.ones(B, N, dtype=torch.bool, device=DEV)
    dh, de_ = torch.randn_like(h), torch.randn_like(e)

    def step():
        hh, ee = h.detach().requires_grad_(True), e.detach().requires_grad_(True)
        blk.flat.grad = None
        h2, e2 = blk(hh, ee, m)
        torch.autograd.backward([h2, e2], [dh, de_])
        return h2.detach(), e2.detach(), hh.grad, ee.grad

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()                                               # eager warm-up (host offset 1)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = step()                                         # host offset 2 is frozen into the graph
    base = blk.rng.offset
    c0 = int(blk._rng_counter.item())
    replays = []
    for r in range(3):
        g.replay()
        torch.cuda.synchronize()
        replays.append([t.clone() for t in out])
    assert int(blk._rng_counter.item()) == c0 + 3
    # the mask acts on the attention weights: h' (and the gradients) change from replay to replay, e' does not depend on it
    assert not torch.equal(replays[0][0], replays[1][0]) and not torch.equal(replays[1][0], replays[2][0])
    for r in range(3):
        blk.rng.offset = base + c0 + r - 1                   # the eager call then uses offset base + c0 + r
        want = step()
        for got, ref in zip(replays[r], want):
            assert torch.equal(got, ref)
