"""Flat parameter arena of the SGSAC agent and its mapping to the reference's state_dict tensors.

One fp32 buffer holds every parameter; the order makes each optimiser's parameter set one contiguous
range (sac.py:60-68, sgsac.py:35-39), so Adam (+ the soft target update) is one launch per optimiser:

    [ Q1 | Q2 | cnn | critic_proj | dec | fdec | actor_proj | actor_mlp ]
      `---------- critic optimiser ---------'
                `----------- aux optimiser -----------'
                                                 `---- actor optimiser (live part) ----'

Internal storage differs from the reference's tensor shapes where the kernels want it (the map is a fixed
permutation, so element-wise Adam / EMA are unaffected and state_dicts round-trip exactly):
  * 3x3 conv weights of SharedCNN layers 2..11 and of the decoder: [Cout][ky][kx][Cin] (reference: [Cout][Cin][ky][kx]);
  * feature maps are NHWC, so the 14112 input columns of the projections and the 14112 output rows of
    decoder.proj are permuted from (c,y,x) to (y,x,c) order;
  * decoder.conv3 is zero-padded from 9 to 16 output channels.
Every tensor starts on a 16-byte boundary (padding floats stay 0: zero grad => Adam leaves them at 0).
"""
from collections import OrderedDict

import torch

FEAT = 32 * 21 * 21
ENC_H = [41, 39, 37, 35, 33, 31, 29, 27, 25, 23, 21]      # spatial size after each SharedCNN conv (84x84 input)
DEC_C3 = 32                                               # stored output channels of decoder.conv3 (9 real; tcgen05 N)


def _al4(n):
    return (n + 3) // 4 * 4


class ParamLayout:
    def __init__(self, action_dim, hidden_dim=1024, proj_dim=100, num_layers=11, num_filters=32, in_ch=9, algorithm="sgsac"):
        assert num_filters == 32 and num_layers == 11 and proj_dim == 100, \
            "the attribution decoder hard-codes 32x21x21 features and a 100-d embedding (modules.py:315-318)"
        self.A, self.H, self.P, self.L, self.nf, self.in_ch = action_dim, hidden_dim, proj_dim, num_layers, num_filters, in_ch
        A, H, P = action_dim, hidden_dim, proj_dim
        ent = OrderedDict()           # canonical name -> (offset, stored numel, canonical shape)
        off = 0

        def add(name, shape, stored=None):
            nonlocal off
            n = 1
            for s in shape:
                n *= s
            st = stored if stored is not None else n
            ent[name] = (off, st, tuple(shape))
            off += _al4(st)

        self.ranges = {}
        start = off
        for q in ("Q1", "Q2"):
            q0 = off
            add(f"{q}.0.weight", (H, P + A)); add(f"{q}.0.bias", (H,))
            add(f"{q}.2.weight", (H, H)); add(f"{q}.2.bias", (H,))
            add(f"{q}.4.weight", (1, H)); add(f"{q}.4.bias", (1,))
            self.ranges[q] = (q0, off)
        self.q_stride = self.ranges["Q2"][0] - self.ranges["Q1"][0]
        enc0 = off
        for i in range(num_layers):
            add(f"cnn.{i}.weight", (num_filters, in_ch if i == 0 else num_filters, 3, 3)); add(f"cnn.{i}.bias", (num_filters,))
        self.ranges["cnn"] = (enc0, off)
        self._add_proj(add, "critic_proj")
        self.ranges["critic_proj"] = (self.ranges["cnn"][1], off)
        self.ranges["critic"] = (start, off)
        self.ranges["critic_q"] = (start, enc0)
        dec0 = off
        add("dec.proj.weight", (FEAT, P + A)); add("dec.proj.bias", (FEAT,))
        add("dec.conv1.weight", (128, 32, 3, 3)); add("dec.conv1.bias", (128,))
        add("dec.conv2.weight", (64, 128, 3, 3)); add("dec.conv2.bias", (64,))
        add("dec.conv3.weight", (9, 64, 3, 3), stored=DEC_C3 * 64 * 9); add("dec.conv3.bias", (9,), stored=DEC_C3)
        self.ranges["dec"] = (dec0, off)
        add("fdec.0.weight", (256, 100)); add("fdec.0.bias", (256,))
        add("fdec.2.weight", (100, 256)); add("fdec.2.bias", (100,))
        if algorithm == "soda":                               # SODAMLP of the SODA encoder + SODAPredictor's SODAMLP (soda.py:19-27)
            soda0 = off
            for pre, fan in (("soda_proj", FEAT), ("soda_pred", P)):
                add(f"{pre}.0.weight", (P, fan)); add(f"{pre}.0.bias", (P,))
                add(f"{pre}.1.weight", (P,)); add(f"{pre}.1.bias", (P,))
                add(f"{pre}.3.weight", (P, P)); add(f"{pre}.3.bias", (P,))
            self.ranges["soda"] = (soda0, off)
        if algorithm == "pad":                                # InverseDynamics (modules.py:284-303) on PAD's own projection (pad.py:18-25)
            self._add_proj(add, "pad_proj")
            add("pad_mlp.0.weight", (H, 2 * P)); add("pad_mlp.0.bias", (H,))
            add("pad_mlp.2.weight", (H, H)); add("pad_mlp.2.bias", (H,))
            add("pad_mlp.4.weight", (A, H)); add("pad_mlp.4.bias", (A,))
        if algorithm == "curl":                               # CURLHead.W (modules.py:264-268): optimised with the critic encoder (curl.py:16-20)
            add("curl.W", (P, P))
        self.ranges["aux"] = (enc0, off)
        act0 = off
        self._add_proj(add, "actor_proj")
        add("actor_mlp.0.weight", (H, P)); add("actor_mlp.0.bias", (H,))
        add("actor_mlp.2.weight", (H, H)); add("actor_mlp.2.bias", (H,))
        add("actor_mlp.4.weight", (2 * A, H)); add("actor_mlp.4.bias", (2 * A,))
        self.ranges["actor"] = (act0, off)
        self.entries = ent
        self.total = off

    def _add_proj(self, add, pre):
        add(f"{pre}.0.weight", (self.P, FEAT)); add(f"{pre}.0.bias", (self.P,))
        add(f"{pre}.1.weight", (self.P,)); add(f"{pre}.1.bias", (self.P,))

    def off(self, name):
        return self.entries[name][0]

    # ---- canonical (reference-shaped) tensor <-> stored flat segment
    def to_stored(self, name, t):
        t = t.detach().to(torch.float32)
        if name.startswith("cnn.") and name.endswith("weight") and not name.startswith("cnn.0."):
            return t.permute(0, 2, 3, 1).reshape(-1)
        if name in ("dec.conv1.weight", "dec.conv2.weight"):
            return t.permute(0, 2, 3, 1).reshape(-1)
        if name == "dec.conv3.weight":
            p = torch.zeros(DEC_C3, 64, 3, 3, dtype=t.dtype, device=t.device)
            p[:9] = t
            return p.permute(0, 2, 3, 1).reshape(-1)
        if name == "dec.conv3.bias":
            p = torch.zeros(DEC_C3, dtype=t.dtype, device=t.device)
            p[:9] = t
            return p
        if name.endswith("proj.0.weight") and name != "dec.proj.weight" and t.shape[-1] == FEAT:
            return t.reshape(t.shape[0], 32, 441).permute(0, 2, 1).reshape(-1)
        if name == "dec.proj.weight":
            return t.reshape(32, 441, t.shape[1]).permute(1, 0, 2).reshape(-1)
        if name == "dec.proj.bias":
            return t.reshape(32, 441).permute(1, 0).reshape(-1)
        return t.reshape(-1)

    def from_stored(self, name, flat):
        _, st, shape = self.entries[name]
        s = flat[:st]
        if name.startswith("cnn.") and name.endswith("weight") and not name.startswith("cnn.0."):
            return s.reshape(shape[0], 3, 3, shape[1]).permute(0, 3, 1, 2).contiguous()
        if name in ("dec.conv1.weight", "dec.conv2.weight"):
            return s.reshape(shape[0], 3, 3, shape[1]).permute(0, 3, 1, 2).contiguous()
        if name == "dec.conv3.weight":
            return s.reshape(DEC_C3, 3, 3, 64).permute(0, 3, 1, 2)[:9].contiguous()
        if name == "dec.conv3.bias":
            return s[:9].clone()
        if name.endswith("proj.0.weight") and name != "dec.proj.weight" and shape[-1] == FEAT:
            return s.reshape(shape[0], 441, 32).permute(0, 2, 1).reshape(shape).contiguous()
        if name == "dec.proj.weight":
            return s.reshape(441, 32, shape[1]).permute(1, 0, 2).reshape(shape).contiguous()
        if name == "dec.proj.bias":
            return s.reshape(441, 32).permute(1, 0).reshape(shape).contiguous()
        return s.reshape(shape).clone()

    def pack(self, params, flat, names=None, base=0):
        """Write canonical tensors into the flat arena (`base` = arena offset of flat[0])."""
        for name in (names if names is not None else self.entries):
            if name not in params:
                continue
            o, st, _ = self.entries[name]
            flat[o - base:o - base + st].copy_(self.to_stored(name, params[name]).to(flat.device))

    def unpack(self, flat, names=None, base=0):
        out = OrderedDict()
        for name in (names if names is not None else self.entries):
            o, st, _ = self.entries[name]
            out[name] = self.from_stored(name, flat[o - base:o - base + _al4(st)])
        return out


# canonical name -> [(module, reference state_dict key)]   (train.py:207-219 saves these three modules)
def reference_key_map(num_layers=11):
    m = OrderedDict()
    for i in range(num_layers):
        for wb in ("weight", "bias"):
            k = f"encoder.shared_cnn.layers.{2 + 2 * i}.{wb}"
            m[f"cnn.{i}.{wb}"] = [("critic", k), ("actor", k), ("attribution_predictor", k)]
    for j in (0, 1):
        for wb in ("weight", "bias"):
            k = f"encoder.projection.projection.{j}.{wb}"
            m[f"critic_proj.{j}.{wb}"] = [("critic", k), ("attribution_predictor", k)]
            m[f"actor_proj.{j}.{wb}"] = [("actor", k)]
    for q in ("Q1", "Q2"):
        for j in (0, 2, 4):
            for wb in ("weight", "bias"):
                m[f"{q}.{j}.{wb}"] = [("critic", f"{q}.trunk.{j}.{wb}")]
    for j in (0, 2, 4):
        for wb in ("weight", "bias"):
            m[f"actor_mlp.{j}.{wb}"] = [("actor", f"mlp.{j}.{wb}")]
    for l in ("proj", "conv1", "conv2", "conv3"):
        for wb in ("weight", "bias"):
            m[f"dec.{l}.{wb}"] = [("attribution_predictor", f"decoder.{l}.{wb}")]
    for j in (0, 2):
        for wb in ("weight", "bias"):
            m[f"fdec.{j}.{wb}"] = [("attribution_predictor", f"features_decoder.{j}.{wb}")]
    m["curl.W"] = [("curl_head", "W")]
    for n in list(m):                                         # curl_head.encoder IS the critic's encoder (curl.py:16)
        if n.startswith(("cnn.", "critic_proj.")):
            m[n] = m[n] + [("curl_head", "encoder." + m[n][0][1][len("encoder."):])]
    for n in list(m):                                         # pad_head.encoder = shared CNN + PAD's own projection (pad.py:18-25)
        if n.startswith("cnn."):
            m[n] = m[n] + [("pad_head", m[n][0][1])]
    for j in (0, 1):
        for wb in ("weight", "bias"):
            m[f"pad_proj.{j}.{wb}"] = [("pad_head", f"encoder.projection.projection.{j}.{wb}")]
    for j in (0, 2, 4):
        for wb in ("weight", "bias"):
            m[f"pad_mlp.{j}.{wb}"] = [("pad_head", f"mlp.{j}.{wb}")]
    for n in list(m):                                         # predictor.encoder = shared CNN + SODAMLP projection (soda.py:19-27)
        if n.startswith("cnn."):
            m[n] = m[n] + [("predictor", m[n][0][1])]
    for j in (0, 1, 3):
        for wb in ("weight", "bias"):
            m[f"soda_proj.{j}.{wb}"] = [("predictor", f"encoder.projection.mlp.{j}.{wb}")]
            m[f"soda_pred.{j}.{wb}"] = [("predictor", f"mlp.mlp.{j}.{wb}")]
    return m
