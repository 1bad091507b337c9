"""ScrimpPolicy — the reference's SCRIMPNet (`net.py:39-155`, `transformer.py:8-100`) restated for batched GPU rollouts.

Same function, same parameters (a reference checkpoint loads through `load_reference_state_dict`), different program:

* The reference "tokenizer" (`net.py:127-134`) is `A = softmax(x @ sum_z wA_z)` over a singleton axis — identically 1 —
  and `T = A @ (x @ sum_z wV_z)`: sixteen copies of one projected feature row.  Here the eight `token_wV` slices are
  summed once per forward (a [512,512] matrix), `token_wA` is kept as a parameter for checkpoint compatibility but
  never enters the graph (its gradient is exactly zero in the reference too).
* Only the class token leaves the transformer (`net.py:141`), so the last block computes queries, the output
  projection and the MLP for that one row instead of all 17 (keys/values still see every token).
* Attention uses `scaled_dot_product_attention` with the reference's scale `dim ** -0.5` (`transformer.py:54`, the
  model width, not the head width).
* `forward` takes `[..., N, C, F, F]` / `[..., N, 4]` with any leading batch shape and has no global `N_AGENTS`.

Dropout (p = 0.2, `net.py:50-51`) is active in `train()` mode exactly as in the reference, which never calls `eval()`;
parity fixtures are generated with `eval()` on both sides (tests/golden/make_ppo_golden.py).
"""
from __future__ import annotations

from typing import Dict, NamedTuple

import torch
import torch.nn as nn
import torch.nn.functional as F


class PolicyOutput(NamedTuple):
    policy: torch.Tensor        # softmax(logits)            [..., N, 5]     net.py:149
    value: torch.Tensor         # state value                [..., N, 1]     net.py:151
    blocking: torch.Tensor      # sigmoid(blocking head)     [..., N, 1]     net.py:153
    policy_sig: torch.Tensor    # sigmoid(logits)            [..., N, 5]     net.py:150
    features: torch.Tensor      # shared features            [..., N, 512]   net.py:147
    logits: torch.Tensor        #                            [..., N, 5]     net.py:148
    cost_value: torch.Tensor    # cost value                 [..., N, 1]     net.py:152


class _Block(nn.Module):
    """Pre-norm transformer block: x + proj(attn(LN(x))), then x + mlp(LN(x))   (transformer.py:8-31, 49-100)."""

    def __init__(self, dim, heads, mlp_dim, dropout):
        super().__init__()
        self.heads, self.dim, self.p = heads, dim, dropout
        self.ln_attn = nn.LayerNorm(dim)
        self.qkv = nn.Linear(dim, 3 * dim)
        self.proj = nn.Linear(dim, dim)
        self.ln_mlp = nn.LayerNorm(dim)
        self.fc1 = nn.Linear(dim, mlp_dim)
        self.fc2 = nn.Linear(mlp_dim, dim)

    def _drop(self, x):
        return F.dropout(x, self.p, self.training)

    def forward(self, x, cls_only: bool):
        b, n, d = x.shape
        h = self.heads
        y = self.ln_attn(x)
        if cls_only:                      # queries for the class token only; keys / values for every token
            q = F.linear(y[:, :1], self.qkv.weight[:d], self.qkv.bias[:d])
            kv = F.linear(y, self.qkv.weight[d:], self.qkv.bias[d:])
            k, v = kv[..., :d], kv[..., d:]
            x = x[:, :1]
        else:
            q, k, v = self.qkv(y).split(d, dim=-1)

        def heads_first(t):              # 'b n (h d) -> b h n d'   (transformer.py:66)
            return t.reshape(b, t.shape[1], h, d // h).transpose(1, 2)
        o = F.scaled_dot_product_attention(heads_first(q), heads_first(k), heads_first(v), scale=float(d) ** -0.5)
        o = o.transpose(1, 2).reshape(b, -1, d)
        x = x + self._drop(self.proj(o))
        y = self.ln_mlp(x)
        y = self._drop(self.fc2(self._drop(F.gelu(self.fc1(y)))))
        return x + y


class ScrimpPolicy(nn.Module):
    def __init__(self, num_channel: int = 6, fov: int = 9, net_size: int = 512, goal_repr: int = 12,
                 vector_len: int = 4, n_actions: int = 5, tokens: int = 16, heads: int = 16, depth: int = 2,
                 mlp_dim: int = 512, dropout: float = 0.2):
        super().__init__()
        if fov != 9:
            raise ValueError("the reference conv/pool stack (net.py:57-66) reduces exactly 9x9 to 1x1")
        q, hlf = net_size // 4, net_size // 2
        self.C, self.F, self.D, self.L, self.p = num_channel, fov, net_size, tokens, dropout
        # observation encoder (net.py:57-66): 9x9 -> pool 4x4 -> 5,6,7 -> pool 3x3 -> 1x1
        self.enc = nn.ModuleDict(dict(
            c1=nn.Conv2d(num_channel, q, 3, 1, 1), c1a=nn.Conv2d(q, q, 3, 1, 1), c1b=nn.Conv2d(q, q, 3, 1, 1),
            c2=nn.Conv2d(q, hlf, 2, 1, 1), c2a=nn.Conv2d(hlf, hlf, 2, 1, 1), c2b=nn.Conv2d(hlf, hlf, 2, 1, 1),
            c3=nn.Conv2d(hlf, net_size - goal_repr, 3, 1, 0)))
        self.goal_fc = nn.Linear(vector_len, goal_repr)                       # net.py:67
        self.mix1 = nn.Linear(net_size, net_size)                             # net.py:68
        self.mix2 = nn.Linear(net_size, net_size)                             # net.py:69
        self.token_wA = nn.Parameter(torch.empty(8, tokens, net_size))        # net.py:72 (inert, see module docstring)
        self.token_wV = nn.Parameter(torch.empty(8, net_size, net_size))      # net.py:74
        self.pos_embedding = nn.Parameter(torch.empty(1, tokens + 1, net_size))
        self.cls_token = nn.Parameter(torch.zeros(1, 1, net_size))
        self.blocks = nn.ModuleList([_Block(net_size, heads, mlp_dim, dropout) for _ in range(depth)])
        self.post = nn.Linear(net_size, net_size)                             # nn_same, applied twice (net.py:145-146)
        self.policy_head = nn.Linear(net_size, n_actions)
        self.value_head = nn.Linear(net_size, 1)
        self.cost_value_head = nn.Linear(net_size, 1)
        self.blocking_head = nn.Linear(net_size, 1)
        self.channels_last = False
        self.reset_parameters()

    def use_channels_last(self, on: bool = True) -> "ScrimpPolicy":
        """NHWC activations for the conv encoder (cuDNN's tensor-core kernels want it: the 9x9 conv stack is ~35 % faster
        in bf16 on B200).  Only the input's memory format is switched — the weights keep their layout, so that the
        learner's flat gradient buffer (and fused Adam) still see dense, identically laid out parameters and gradients."""
        self.channels_last = on
        return self

    # ---- initialisation (net.py:11-36, 72-99; transformer.py:29-62): same distributions, own RNG order -----------
    def reset_parameters(self):
        def glorot(w, fan_in, fan_out):
            bound = (6.0 / (fan_in + fan_out)) ** 0.5
            nn.init.uniform_(w, -bound, bound)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):                                       # weights_init 'Conv' branch
                o, i, kh, kw = m.weight.shape
                glorot(m.weight, i * kh * kw, kh * kw * o)
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Linear):                                     # weights_init 'Linear' branch
                glorot(m.weight, m.weight.shape[1], m.weight.shape[0])
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.LayerNorm):
                nn.init.ones_(m.weight); nn.init.zeros_(m.bias)
        nn.init.xavier_uniform_(self.token_wA)
        nn.init.xavier_uniform_(self.token_wV)
        nn.init.normal_(self.pos_embedding, std=0.02)
        nn.init.zeros_(self.cls_token)

    # ---- forward -----------------------------------------------------------------------------------------------
    def features(self, obs: torch.Tensor, vector: torch.Tensor) -> torch.Tensor:
        """[R, C, F, F], [R, 4] -> [R, 512]   (net.py:105-146)."""
        e = self.enc
        if self.channels_last:
            obs = obs.contiguous(memory_format=torch.channels_last)
        if obs.is_cuda and self.channels_last and not torch.is_grad_enabled():
            # rollout / evaluation path: cuDNN's fused convolution + bias + ReLU epilogue (one pass over the activations
            # instead of two per layer; identical values).  It has no autograd formula, hence no_grad only.
            dt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else obs.dtype
            x = obs.to(dt)

            def cr(x, m):
                return torch.cudnn_convolution_relu(x, m.weight.to(dt), m.bias.to(dt), m.stride, m.padding, m.dilation, 1)
            x = cr(cr(cr(x, e["c1"]), e["c1a"]), e["c1b"])
            x = F.max_pool2d(x, 2)
            x = cr(cr(cr(x, e["c2"]), e["c2a"]), e["c2b"])
            x = F.max_pool2d(x, 2)
            x = cr(x, e["c3"]).flatten(1)
        else:
            x = F.relu(e["c1"](obs)); x = F.relu(e["c1a"](x)); x = F.relu(e["c1b"](x))
            x = F.max_pool2d(x, 2)
            x = F.relu(e["c2"](x)); x = F.relu(e["c2a"](x)); x = F.relu(e["c2b"](x))
            x = F.max_pool2d(x, 2)
            x = F.relu(e["c3"](x).flatten(1))
        g = F.relu(self.goal_fc(vector))
        x3 = torch.cat((x, g), dim=-1)
        h = self.mix2(F.relu(self.mix1(x3)))
        h = F.relu(h + x3)
        tok = h @ self.token_wV.sum(dim=0).to(h.dtype)                         # every token row (net.py:127-134)
        R = h.shape[0]
        seq = torch.cat((self.cls_token.to(tok.dtype).expand(R, 1, -1), tok.unsqueeze(1).expand(R, self.L, -1)), dim=1)
        seq = F.dropout(seq + self.pos_embedding.to(tok.dtype), self.p, self.training)
        last = len(self.blocks) - 1
        for k, blk in enumerate(self.blocks):
            seq = blk(seq, cls_only=(k == last))
        y = seq[:, 0]
        return self.post(self.post(y))

    def forward(self, obs: torch.Tensor, vector: torch.Tensor) -> PolicyOutput:
        lead = obs.shape[:-3]                                                  # (..., N)
        f = self.features(obs.reshape(-1, self.C, self.F, self.F), vector.reshape(-1, vector.shape[-1]))
        f = f.reshape(*lead, self.D)
        logits = self.policy_head(f)
        return PolicyOutput(policy=F.softmax(logits, dim=-1), value=self.value_head(f),
                            blocking=torch.sigmoid(self.blocking_head(f)), policy_sig=torch.sigmoid(logits),
                            features=f, logits=logits, cost_value=self.cost_value_head(f))

    # ---- reference checkpoints ----------------------------------------------------------------------------------
    def load_reference_state_dict(self, ref_sd: Dict[str, torch.Tensor]) -> None:
        """Loads `SCRIMPNet.state_dict()` (e.g. `net_checkpoint.pkl['model']`, driver.py:189-194)."""
        own = self.state_dict()
        missing = [r for r in REFERENCE_KEY_MAP.values() if r not in ref_sd]
        if missing:
            raise KeyError(f"reference state dict lacks {missing[:4]}...")
        for mine, ref in REFERENCE_KEY_MAP.items():
            if own[mine].shape != ref_sd[ref].shape:
                raise ValueError(f"{ref}: shape {tuple(ref_sd[ref].shape)} != {tuple(own[mine].shape)}")
            own[mine].copy_(ref_sd[ref])

    def reference_state_dict(self) -> Dict[str, torch.Tensor]:
        own = self.state_dict()
        return {ref: own[mine].clone() for mine, ref in REFERENCE_KEY_MAP.items()}


def _key_map(depth: int = 2) -> Dict[str, str]:
    m = {}
    for mine, ref in (("enc.c1", "conv1"), ("enc.c1a", "conv1a"), ("enc.c1b", "conv1b"), ("enc.c2", "conv2"),
                      ("enc.c2a", "conv2a"), ("enc.c2b", "conv2b"), ("enc.c3", "conv3"),
                      ("goal_fc", "fully_connected_1"), ("mix1", "fully_connected_2"), ("mix2", "fully_connected_3"),
                      ("post", "nn_same"), ("policy_head", "policy_layer"), ("value_head", "value_layer"),
                      ("cost_value_head", "cost_value_layer"), ("blocking_head", "blocking_layer")):
        for s in ("weight", "bias"):
            m[f"{mine}.{s}"] = f"{ref}.{s}"
    for n in ("token_wA", "token_wV", "pos_embedding", "cls_token"):
        m[n] = n
    for k in range(depth):
        a, f = f"transformer.layers.{k}.0.fn", f"transformer.layers.{k}.1.fn"      # Residual(LayerNormalize(...))
        for mine, ref in ((f"blocks.{k}.ln_attn", f"{a}.norm"), (f"blocks.{k}.qkv", f"{a}.fn.to_qkv"),
                          (f"blocks.{k}.proj", f"{a}.fn.nn1"), (f"blocks.{k}.ln_mlp", f"{f}.norm"),
                          (f"blocks.{k}.fc1", f"{f}.fn.nn1"), (f"blocks.{k}.fc2", f"{f}.fn.nn2")):
            for s in ("weight", "bias"):
                m[f"{mine}.{s}"] = f"{ref}.{s}"
    return m


REFERENCE_KEY_MAP = _key_map()
