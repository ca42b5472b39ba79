import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from tests.conftest import load_golden, max_norm_err
from grit_b200 import MSDeformAttn
for name in ("module_ref2", "module_ref4"):
    g = load_golden(name)
    mod = MSDeformAttn(int(g["d_model"]), int(g["n_levels"]), int(g["n_heads"]), int(g["n_points"]))
    mod.load_state_dict({k[len("param."):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param.")})
    mod = mod.to("cuda", torch.float64)
    cu = lambda k: torch.from_numpy(g[k]).to("cuda")
    query = cu("query").requires_grad_(True); src = cu("input_flatten").requires_grad_(True)
    out = mod(query, cu("reference_points"), src, cu("shapes"), cu("level_start"), cu("padding_mask"))
    out.backward(cu("grad_out"))
    print(name, "out", max_norm_err(out.detach().cpu().numpy(), g["out"]), "gq", max_norm_err(query.grad.cpu().numpy(), g["grad_query"]),
          "gsrc", max_norm_err(src.grad.cpu().numpy(), g["grad_input_flatten"]))
    for k, p in mod.named_parameters():
        print("   ", k, max_norm_err(p.grad.cpu().numpy(), g["grad." + k]), float(np.abs(g["grad."+k]).max()))
