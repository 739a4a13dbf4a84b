"""Development probe: iteration-0 (RTN) loss of the Mixtral fixture block -- product (grouped path) vs the oracle loop on
the GPU vs the oracle loop on the CPU, plus the routing decisions each side takes."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_moe import _mixtral_block  # noqa: E402

from auto_round_b200.quantizer import SignRoundQuantizer  # noqa: E402
from auto_round_b200.schemes import parse_scheme  # noqa: E402
from oracle import signround as S  # noqa: E402

DEV = torch.device("cuda", 0)
rec = torch.load(os.path.join(ROOT, "tests", "golden", "block_mixtral_mxfp4.pt"), weights_only=False)
b = rec["blocks"][0]
osc = S.LayerScheme(4, 32, True, "mx_fp")
masks = [(ids != -100).to(torch.long) for ids in b["input_ids"]]
batch = [0, 1, 2, 3]
routes = {}


def hook(tag):
    def f(mod, args, out):
        routes[tag] = out[2].detach().cpu().clone()
    return f


def to_dev(o, d):
    if isinstance(o, torch.Tensor):
        return o.to(d)
    if isinstance(o, (list, tuple)):
        return type(o)(to_dev(x, d) for x in o)
    return o


for tag, dev in (("oracle_cpu", "cpu"), ("oracle_gpu", DEV)):
    blk = _mixtral_block(b["block_state"], False, dev)
    blk.block_sparse_moe.gate.register_forward_hook(hook(tag)) if hasattr(blk, "block_sparse_moe") else blk.mlp.gate.register_forward_hook(hook(tag))
    others = {k: to_dev(v, dev) for k, v in b["others"].items()}
    res = S.tune_block(blk, [t.to(dev) for t in b["inputs"]], others, [t.to(dev) for t in b["fp_outputs"]], lambda n, m: osc, iters=1,
                       batch_size=4, token_masks=[m.to(dev) for m in masks], sampler=S.ReplaySampler([batch]))
    print(tag, "loss0 =", res.losses[0])
blk = _mixtral_block(b["block_state"], True, DEV)
for p in blk.parameters():
    p.requires_grad_(False)
moe = blk.block_sparse_moe if hasattr(blk, "block_sparse_moe") else blk.mlp
moe.gate.register_forward_hook(hook("ours"))
for graph in (False,):
    q = SignRoundQuantizer(parse_scheme("MXFP4", {"act_bits": 16}), iters=1, batch_size=4, use_cuda_graph=graph)
    q.quantize_block(blk, [t.to(DEV) for t in b["inputs"]], b["others"], [t.to(DEV) for t in b["fp_outputs"]], None, None,
                     input_ids=b["input_ids"], sampler=S.ReplaySampler([batch]))
    print("ours loss0 =", q.last_result.losses[0])
for a in ("oracle_gpu", "ours"):
    if a in routes and "oracle_cpu" in routes:
        x, y = routes[a].sort(-1).values, routes["oracle_cpu"].sort(-1).values
        print(a, "tokens routed differently from oracle_cpu:", int((x != y).any(-1).sum()), "of", x.shape[0])
