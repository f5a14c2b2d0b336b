"""CPU emulation of the tensor-core mode's rounding points (diagnostic; which rounding owns the error tail).

Runs the oracle's `cone_forward` arithmetic on the windows of the end-to-end test case with fp16 rounding inserted at
the places the CUDA pipeline rounds (transformer_core in csrc/api.cu) and prints the error of spans / foreground
probabilities against the plain fp32 oracle, for a list of rounding configurations.

    python profiles/tc_emulate.py [config ...]
"""
import os
import sys
import math

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cone_b200.config import EGO4D, MAD768  # noqa: E402
from cone_b200.synth import make_dataset  # noqa: E402
from cone_b200.weights import init_state_dict  # noqa: E402
from oracle import cone_oracle as O  # noqa: E402


def h(x):  # fp16 round trip
    return x.half().float()


def hl(x):  # hi + lo fp16 split (the split GEMM's operand accuracy)
    hi = x.half().float()
    return hi + (x - hi).half().float()


class Emu:
    """flags: set of rounding points that are ON (fp16).  Names:
    src        window rows (operand of layer-0 QKV, first residual)
    w_enc      encoder weights            w_dec   decoder weights
    qkv        q|k|v GEMM outputs         p       attention probabilities     att   attention output
    ln_op      LayerNorm outputs as GEMM operands          ln_res   LayerNorm outputs as residuals
    hid        FFN hidden
    d_op       decoder GEMM operands (tgt16, dqkin16, datt16, dh16)    d_qkv  decoder q|k|v, dqt16   d_p   decoder probabilities
    d_pm       pooled memory (dpm16)      mem    encoder output read by the cross attention
    """

    def __init__(self, sd, cfg, flags, split=()):
        self.sd, self.cfg, self.f, self.split = sd, cfg, set(flags), set(split)

    def r(self, name, x):
        if name in self.split:
            return hl(x)
        return h(x) if name in self.f else x

    def W(self, key, which):
        w = self.sd[key]
        if which in self.split:
            return hl(w)
        return h(w) if which in self.f else w

    def mha_enc(self, p, src_op, pos, pad_mask):
        sd, H = self.sd, self.cfg.nheads
        E = src_op.shape[-1]
        hd = E // H
        w, b = self.W(p + ".in_proj_weight", "w_enc"), sd[p + ".in_proj_bias"]
        wf = sd[p + ".in_proj_weight"]
        # q = (src W^T + b) [fp16] + pos W^T [fp16 table]
        qkv = self.r("qkv", F.linear(src_op, w, b))
        posq = self.r("qkv", F.linear(pos, wf[:E]))
        posk = self.r("qkv", F.linear(pos, wf[E:2 * E]))
        q = self.r("qkv", qkv[..., :E] + posq)
        k = self.r("qkv", qkv[..., E:2 * E] + posk)
        v = qkv[..., 2 * E:]
        B, S, _ = q.shape
        q = q.reshape(B, S, H, hd).transpose(1, 2) * math.sqrt(1.0 / hd)
        k = k.reshape(B, S, H, hd).transpose(1, 2)
        v = v.reshape(B, S, H, hd).transpose(1, 2)
        s = q @ k.transpose(-1, -2)
        s = s.masked_fill(pad_mask[:, None, None, :], float("-inf"))
        m = s.max(-1, keepdim=True).values
        if "p_f16exp" in self.f:  # ex2.approx.f16x2: fp16 argument, fp16 result
            x16 = h((s - m) * 1.4426950408889634)
            e = torch.exp2(x16)
        else:
            e = torch.exp(s - m)
        pe = self.r("p", e)
        o = (pe @ v) / pe.sum(-1, keepdim=True)  # row sums from the rounded P (ones-column MMA)
        o = o.transpose(1, 2).reshape(B, S, E)
        return self.r("att", o)

    def enc_layer(self, p, src_op, src_res, pos, pad_mask):
        sd = self.sd
        att = self.mha_enc(p + ".self_attn", src_op, pos, pad_mask)
        x = F.linear(att, self.W(p + ".self_attn.out_proj.weight", "w_enc"), sd[p + ".self_attn.out_proj.bias"]) + src_res
        x = O._ln(sd, p + ".norm1", x)
        x_op, x_res = self.r("ln_op", x), self.r("ln1_res", x)
        hid = self.r("hid", F.relu(F.linear(x_op, self.W(p + ".linear1.weight", "w_ffn"), sd[p + ".linear1.bias"])))
        y = F.linear(hid, self.W(p + ".linear2.weight", "w_ffn"), sd[p + ".linear2.bias"]) + x_res
        y = O._ln(sd, p + ".norm2", y)
        return self.r("ln_op", y), self.r("ln2_res", y)

    def mha_plain(self, p, q_in, k_in, v_in, pad_mask, wname, opname, pname, qname="d_qkv"):
        """decoder attention with operand roundings (formulation-independent approximation)"""
        sd, H = self.sd, self.cfg.nheads
        E = q_in.shape[-1]
        hd = E // H
        w, b = self.W(p + ".in_proj_weight", wname), sd[p + ".in_proj_bias"]
        q = self.r(qname, F.linear(q_in, w[:E], b[:E]))
        k = self.r("d_qkv" if pad_mask is None else "none", F.linear(k_in, w[E:2 * E], b[E:2 * E]))
        v = self.r("d_qkv" if pad_mask is None else "none", F.linear(v_in, w[2 * E:], b[2 * E:]))
        B, Lq, _ = q.shape
        Lk = k.shape[1]
        q = q.reshape(B, Lq, H, hd).transpose(1, 2) * math.sqrt(1.0 / hd)
        k = k.reshape(B, Lk, H, hd).transpose(1, 2)
        v = v.reshape(B, Lk, H, hd).transpose(1, 2)
        s = q @ k.transpose(-1, -2)
        if pad_mask is not None:
            s = s.masked_fill(pad_mask[:, None, None, :], float("-inf"))
        m = s.max(-1, keepdim=True).values
        e = self.r(pname, torch.exp(s - m))
        o = (e @ v) / e.sum(-1, keepdim=True)
        o = self.r(opname, o.transpose(1, 2).reshape(B, Lq, E))
        return F.linear(o, self.W(p + ".out_proj.weight", wname), sd[p + ".out_proj.bias"])

    def forward(self, src_txt, txt_mask, src_vid, vid_mask):
        sd, cfg = self.sd, self.cfg
        v = O.input_proj(sd, "input_vid_proj", src_vid, cfg.n_input_proj)
        t = O.input_proj(sd, "input_txt_proj", src_txt, cfg.n_input_proj)
        src = torch.cat([v, t], 1)
        mask = torch.cat([vid_mask, txt_mask], 1).bool()
        pad = ~mask
        d = src.shape[-1]
        pos = torch.cat([O.position_embedding_sine(vid_mask, d), torch.zeros_like(t)], 1)
        op, res = self.r("src", src), self.r("src_res", src)
        for l in range(cfg.enc_layers):
            op, res = self.enc_layer(f"transformer.encoder.layers.{l}", op, res, pos, pad)
        mem_exact = res
        mem = self.r("mem", op)
        B = src.shape[0]
        qpos = sd["query_embed.weight"][None].repeat(B, 1, 1)
        tgt = torch.zeros_like(qpos)
        hs = None
        for l in range(cfg.dec_layers):
            p = f"transformer.decoder.layers.{l}"
            qk = self.r("d_in", tgt + qpos)
            t2 = self.mha_plain(p + ".self_attn", qk, qk, self.r("d_in", tgt), None, "w_dec", "d_att", "d_p")
            tgt = O._ln(sd, p + ".norm1", tgt + t2)
            t2 = self.mha_plain(p + ".multihead_attn", self.r("d_inx", tgt + qpos), mem + pos, mem, pad, "w_decx", "d_pm", "d_p", "d_qx")
            tgt = O._ln(sd, p + ".norm2", tgt + t2)
            hid = self.r("d_hid", F.relu(F.linear(self.r("d_fin", tgt), self.W(p + ".linear1.weight", "w_dffn"), sd[p + ".linear1.bias"])))
            t2 = F.linear(hid, self.W(p + ".linear2.weight", "w_dffn"), sd[p + ".linear2.bias"])
            tgt = O._ln(sd, p + ".norm3", tgt + t2)
            hs = O._ln(sd, "transformer.decoder.norm", tgt)
        logits = F.linear(hs, sd["class_embed.weight"], sd["class_embed.bias"])
        spans = O.mlp(sd, "span_embed", hs, 3).sigmoid()
        return logits, spans


ALL = {"src", "src_res", "w_enc", "w_ffn", "qkv", "p", "att", "ln_op", "ln1_res", "ln2_res", "hid", "w_dec", "w_decx", "w_dffn",
       "d_in", "d_inx", "d_att", "d_hid", "d_fin", "d_qkv", "d_qx", "d_p", "d_pm", "mem"}


def batches(sd, cfg, ds):
    """the eval batches of the oracle pipeline (windows of every query, padded per batch)"""
    out = []
    with torch.no_grad():
        ctx = [O.stage0_video_context(sd, torch.from_numpy(O.l2_normalize_np(v))) for v in ds.videos]
        raw = [torch.from_numpy(v) for v in ds.videos]
        qs = ds.queries
        for b0 in range(0, len(qs), cfg.eval_bsz):
            vids, toks = [], []
            for q in qs[b0:b0 + cfg.eval_bsz]:
                tok, _ = O.prepare_query_text(q.tokens, q.cls, cfg.max_q_l)
                rl, _ = O.stage1_ranklist(ctx[q.video_idx], torch.from_numpy(O.l2_normalize_np(q.cls)), cfg.max_v_l)
                for (s, n, rows) in O.slice_query_windows(raw[q.video_idx], rl, cfg.topk_window, cfg.max_v_l):
                    vids.append(rows)
                    toks.append(tok)
            out.append(O.pad_sequences(toks) + O.pad_sequences(vids))
    return out


def run(configs, seeds=((21, 33), (5, 7), (9, 11)), cfg=None, mad=False):
    cfg = cfg or EGO4D.replace(eval_bsz=8)
    cases = []
    for ws, dsd in seeds:
        sd = init_state_dict(cfg, ws)
        ds = make_dataset(cfg, 1, [45000], 64, seed=dsd) if mad else make_dataset(cfg, 4, [900, 455, 91, 1300], 4, seed=dsd)
        bt = batches(sd, cfg, ds)
        with torch.no_grad():
            ref = [Emu(sd, cfg, set()).forward(*b) for b in bt]
        cases.append((sd, bt, ref))
    for name, flags, split in configs:
        line = []
        for sd, bt, ref in cases:
            errs = []
            with torch.no_grad():
                for b, (rl, rs) in zip(bt, ref):
                    lg, sp = Emu(sd, cfg, flags, split).forward(*b)
                    errs.append((sp - rs).abs().ravel())
                    errs.append((F.softmax(lg, -1)[..., 0] - F.softmax(rl, -1)[..., 0]).abs().ravel())
            e = torch.cat(errs).numpy()
            line.append(f"rms {np.sqrt(np.mean(e ** 2)):.2e} p99 {np.percentile(e, 99):.2e} max {e.max():.2e} n>1e-3 {int((e > 1e-3).sum())}")
        print(f"{name:34s} | " + " | ".join(line), flush=True)


if __name__ == "__main__":
    torch.set_num_threads(8)
    RES = {"ln1_res", "ln2_res", "src_res"}
    DEC = {"w_dec", "d_op", "d_qkv"}
    DOP = {"d_in", "d_inx", "d_att", "d_fin"}
    NOW = ALL - RES - {"w_dec", "w_decx", "w_dffn", "d_qkv"} - DOP      # the chain as built (decoder hidden fp16)
    cf = [("as built", NOW, ()),
          ("as built + f16x2 exp in the encoder attention", NOW | {"p_f16exp"}, ()),
          ]
    if len(sys.argv) > 1 and sys.argv[1] == "ablate":
        cf = [("all fp16 (current pipeline)", ALL, ())]
        for k in sorted(ALL):
            cf.append((f"all but {k}", ALL - {k}, ()))
        for k in sorted(ALL):
            cf.append((f"only {k}", {k}, ()))
    if len(sys.argv) > 1 and sys.argv[1] == "mad":
        run(cf, seeds=((5, 11),), cfg=MAD768, mad=True)
    else:
        run(cf)
