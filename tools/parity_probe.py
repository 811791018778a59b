"""Block-by-block comparison of the frozen-tower engine with the device-arithmetic oracle (diagnostic)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import clipself_oracle as O, device_arith_oracle as DA
from clipself_b200.tower import TowerEngine, TowerCfg
from clipself_b200 import ops, _lib as L

which = sys.argv[1] if len(sys.argv) > 1 else "tiny"
ocfg = O.CFG_TINY if which == "tiny" else O.CFG_B16
cfg = TowerCfg(image_size=ocfg.image_size, patch=ocfg.patch, width=ocfg.width, heads=ocfg.heads, layers=ocfg.layers,
               hidden=ocfg.hidden, embed_dim=ocfg.embed_dim, pt_seq_len=ocfg.pt_seq_len, ln_eps=ocfg.ln_eps)
dev = torch.device("cuda")
sd = O.synth_tower_weights(ocfg, 11)
images, boxes, crops = O.synth_batch(ocfg, 2, 4, 13, kind="grid")
imgs = crops.flatten(0, 1)
eng = TowerEngine(cfg, sd, dev)
eng.use_graphs = False
n = imgs.shape[0]
ws = eng.workspace(n)
M = n * cfg.tokens
rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()
ar = DA.Arith(False)
cos, sin = O.rope_tables(cfg.grid, cfg.head_dim, cfg.pt_seq_len)
with torch.no_grad():
    x_ref = DA.embed_tokens(ar, sd, imgs, ocfg)
    eng.embed(imgs.to(dev), ws.x, ws)
    print("embed x", rel(ws.x[:M].cpu().view_as(x_ref), x_ref), " xb exact:", torch.equal(ws.xb[:M].cpu().float().view_as(x_ref), x_ref.to(torch.bfloat16).float()))
    for i in range(cfg.layers):
        # oracle pieces of block i
        p = f"blocks.{i}."
        xb = ar.r(x_ref)
        Wqkv = torch.cat([sd[p + "attn.q_proj.weight"], sd[p + "attn.k_proj.weight"], sd[p + "attn.v_proj.weight"]])
        bqkv = torch.cat([sd[p + "attn.q_bias"], torch.zeros_like(sd[p + "attn.q_bias"]), sd[p + "attn.v_bias"]])
        qkv_pre = DA._folded_linear(ar, x_ref, xb, Wqkv, sd[p + "norm1.weight"], sd[p + "norm1.bias"], bqkv, cfg.ln_eps)
        att_ref = DA._attention_core(ar, qkv_pre, ocfg, cos, sin)
        x_ref = DA.block_fold(ar, x_ref, sd, i, ocfg, cos, sin)
        # device: run block, look at intermediates left in the workspace
        eng.block_inplace(i, ws, n)
        torch.cuda.synchronize()
        B, N, D = n, cfg.tokens, cfg.width
        H, hd = cfg.heads, 64
        q, k, v = qkv_pre.reshape(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)
        q = torch.cat([q[:, :, :1], O.rope_apply(q[:, :, 1:], cos, sin)], dim=2)
        k = torch.cat([k[:, :, :1], O.rope_apply(k[:, :, 1:], cos, sin)], dim=2)
        qkv_ref = torch.stack([q, k, v]).permute(1, 3, 0, 2, 4).reshape(M, 3 * D)
        dq = ws.qkv[:M].cpu().float()
        print(f"block {i}: qkv rel {rel(dq, qkv_ref):.3e} (bf16-rounded ref {rel(dq, qkv_ref.to(torch.bfloat16).float()):.3e})"
              f"  att rel {rel(ws.att[:M].cpu().float().view_as(att_ref), att_ref):.3e}"
              f"  x rel {rel(ws.x[:M].cpu().view_as(x_ref), x_ref):.3e}")
