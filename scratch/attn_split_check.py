"""split-KV tail vs unsplit kernel (bitwise-close) and vs torch fp32 on several shapes incl. ragged T"""
import sys, torch
sys.path.insert(0, "/root/repo")
from signerf_b200 import nn_ops, _lib
for shape in (0, 1):
  _lib.set_option('attn_shape', shape)
  for (B, heads, T) in ((2, 20, 4096), (1, 3, 1000), (2, 10, 16384), (1, 2, 777), (2, 5, 2304)):
      C = heads * 64
      g = torch.Generator(device="cuda").manual_seed(T)
      q, k, v = (torch.randn(B * T, C, device="cuda", generator=g).half() for _ in range(3))
      _lib.set_option("attn_split", 0); a = nn_ops.attention_f16(q, k, v, B, heads).float()
      _lib.set_option("attn_split", 1); b = nn_ops.attention_f16(q, k, v, B, heads).float()
      need = _lib.load().sgn_attention_workspace_bytes(B, heads, T, T)
      msg = f"shape {shape} B{B} h{heads} T{T}: ws {need} B, split vs unsplit max diff {(a - b).abs().max().item():.2e}"
      if T <= 4096:
          qf, kf, vf = (t.float().view(B, T, heads, 64).transpose(1, 2) for t in (q, k, v))
          ref = (torch.softmax(qf @ kf.transpose(-1, -2) / 8.0, -1) @ vf).transpose(1, 2).reshape(B * T, C)
          msg += f"  rel-L2 vs torch {((b - ref).norm() / ref.norm()).item():.2e}"
      print(msg, flush=True)
