"""Host-side layout of EGTStack (no GPU): ONE flat parameter whose slices are the layers' weights, in execution
order, 16-byte aligned; weights load by the reference's layer names."""
import torch


def test_stack_owns_one_flat_parameter_with_aliasing_views():
    import egt_b200
    L = 3
    st = egt_b200.EGTStack(L, ffn=True, model_width=64, edge_width=8, num_heads=8)
    params = list(st.parameters())
    assert len(params) == 1 and params[0] is st.flat                       # one buffer = one all-reduce
    assert len(st._owners) == 3 * L                                        # block, node FFN, edge FFN per layer
    end = 0
    for m, (off, n) in zip(st._owners, st._spans):
        assert off % 4 == 0 and off >= end and n == m.flat.numel()
        end = off + n
        assert m.flat.data_ptr() == st.flat.data_ptr() + 4 * off           # a view, not a copy
    assert end <= st.flat.numel()
    # writing through a layer's named view changes the stack's parameter
    blk = st.layers[1].block
    with torch.no_grad():
        blk.view('dense_qkv_bias').fill_(3.5)
    off, _ = st._spans[3]
    o, shape = blk.layout['dense_qkv_bias']
    assert float(st.flat.data[off + o]) == 3.5 and float(st.flat.data[off + o + shape[0] - 1]) == 3.5


def test_stack_loads_keras_weights_by_layer_tag():
    import egt_b200
    st = egt_b200.EGTStack(2, ffn=False, model_width=16, edge_width=8, num_heads=4)
    w = {}
    for i, layer in enumerate(st.layers):
        for k, v in layer.keras_weights().items():
            w[k] = torch.full_like(v, float(i + 1))
    st.load_keras_weights(w)
    for i, layer in enumerate(st.layers):
        assert layer.tag == f'{i:0>2d}'
        assert float(layer.view('dense_mha_kernel').min()) == float(i + 1) == float(layer.view('dense_mha_kernel').max())
