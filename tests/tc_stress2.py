"""Debug harness: fused FeatureCross fwd/bwd on the tcgen05 engine vs torch float64 autograd, many seeds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import keras_rs_b200 as K

K.set_gemm_engine(os.environ.get("ENGINE", "tcgen05"))
B, D = 640, 832
R = int(os.environ.get("REPS", "20"))
for P in (64, None):
    for act in (None, "relu"):
        worst = {}
        bad = {}
        for r in range(R):
            g = torch.Generator(device="cuda").manual_seed(1000 + r)
            x0 = torch.randn((B, D), device="cuda", generator=g)
            x = torch.randn((B, D), device="cuda", generator=g)
            gy = torch.randn((B, D), device="cuda", generator=g)
            layer = K.layers.FeatureCross(projection_dim=P, diag_scale=0.25, pre_activation=act,
                                          bias_initializer=K.initializers.RandomUniform(-0.5, 0.5, seed=r))
            tx0, tx = x0.clone().requires_grad_(True), x.clone().requires_grad_(True)
            y = layer(tx0, tx)
            y.backward(gy)
            # float64 reference with the SAME relu mask as the kernel's fp32 pre-activation
            d = lambda t: None if t is None else t.detach().double().requires_grad_(True)
            rx0, rx, V, b = d(x0), d(x), d(layer.kernel), d(layer.bias)
            U = d(layer.down_proj_kernel) if P is not None else None
            h = rx if U is None else rx @ U
            z = h @ V + b
            if act == "relu":
                hz = (x if P is None else K.ops.linear_no_bias(x, layer.down_proj_kernel.detach()))
                z32 = K.ops.dense(hz.contiguous(), layer.kernel.detach(), layer.bias.detach(), 0)
                a = z * (z32.double() > 0)
            else:
                a = z
            ry = rx0 * (a + 0.25 * rx) + rx
            ry.backward(gy.double())
            pairs = dict(y=(y, ry), dx0=(tx0.grad, rx0.grad), dx=(tx.grad, rx.grad), dV=(layer.kernel.grad, V.grad), db=(layer.bias.grad, b.grad))
            if P is not None:
                pairs["dU"] = (layer.down_proj_kernel.grad, U.grad)
            for k, (got, ref) in pairs.items():
                e = float((got.double() - ref.detach()).abs().max() / ref.detach().abs().max())
                worst[k] = max(worst.get(k, 0.0), e)
                bad[k] = bad.get(k, 0) + (e > 1e-5)
        print(f"P={P} act={act}: " + "  ".join(f"{k}:{worst[k]:.1e}({bad[k]})" for k in worst), flush=True)
