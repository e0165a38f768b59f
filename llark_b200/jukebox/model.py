"""B200-native Jukebox "5b" front-end modules: the objects that jukebox/main.py receives from
`jukebox.make_models.make_vqvae / make_prior` in the reference (jukebox/main.py:176-200), re-implemented on the
C-ABI kernels (conv stack, VQ, LayerNorm, factored attention, tcgen05 GEMMs, pool).

Only what the reference's call sites touch is provided:
    vqvae.encode(x[N,T,1]) -> list of code tensors, last = top level                (jukebox/main.py:61-63)
    top_prior.raw_to_tokens, .labeller.get_batch_labels, .get_y, .get_cond           (jukebox/main.py:74-93)
    top_prior.prior.only_encode, top_prior.prior.forward(x, x_cond=, y_cond=, encoder_kv=None, fp16=False)
                                                                                      (jukebox/main.py:101-110)
Architecture: SURVEY Appendix A (openai/jukebox@08efbbc, not vendored). Levels 0/1 of the VQ-VAE are dead work
in the reference (only zs[-1] is kept, :63) and are skipped: results are identical.

Weights: no checkpoint exists offline; `state` dicts use the names of oracle/jukebox_oracle.py so parity tests can
load identical parameters. Conv1D weights of the prior are stored TRANSPOSED ([n_out, n_in], K-major) once at
load, so every GEMM reads its operands K-major (TF32 UMMA supports K-major only).
"""
import math
from dataclasses import dataclass
from types import SimpleNamespace

import torch

from llark_b200 import ops

BF16 = torch.bfloat16


@dataclass
class JukeboxConfig:
    sample_length: int = 1048576
    downs_t: tuple = (3, 2, 2)
    width: int = 32
    depth: int = 4
    emb_width: int = 64
    l_bins: int = 2048
    dilation_growth_rate: int = 3
    n_ctx: int = 8192
    prior_width: int = 4800
    prior_depth: int = 36
    heads: int = 8
    m_attn: float = 0.25
    m_mlp: float = 1.0
    blocks: int = 128
    ln_eps: float = 1e-5

    @property
    def hop(self):
        return 2 ** sum(self.downs_t)

    @property
    def attn_width(self):
        return int(self.m_attn * self.prior_width)

    @property
    def mlp_width(self):
        return int(self.m_mlp * self.prior_width)

    @property
    def block_ctx(self):
        return self.n_ctx // self.blocks


def random_vqvae_state(cfg, seed=0, device="cuda"):
    g = torch.Generator().manual_seed(seed)
    st = {}
    c_in = 1
    for b, down_t in enumerate(cfg.downs_t):
        for i in range(down_t):
            cin = c_in if i == 0 else cfg.width
            st[f"enc.{b}.down.{i}.w"] = torch.randn(cfg.width, cin, 4, generator=g) / math.sqrt(cin * 4)
            st[f"enc.{b}.down.{i}.b"] = torch.randn(cfg.width, generator=g) * 0.02
            for d in range(cfg.depth):
                st[f"enc.{b}.res.{i}.{d}.w1"] = torch.randn(cfg.width, cfg.width, 3, generator=g) / math.sqrt(cfg.width * 3)
                st[f"enc.{b}.res.{i}.{d}.b1"] = torch.randn(cfg.width, generator=g) * 0.02
                st[f"enc.{b}.res.{i}.{d}.w2"] = torch.randn(cfg.width, cfg.width, 1, generator=g) / math.sqrt(cfg.width)
                st[f"enc.{b}.res.{i}.{d}.b2"] = torch.randn(cfg.width, generator=g) * 0.02
        st[f"enc.{b}.proj.w"] = torch.randn(cfg.emb_width, cfg.width, 3, generator=g) / math.sqrt(cfg.width * 3)
        st[f"enc.{b}.proj.b"] = torch.randn(cfg.emb_width, generator=g) * 0.02
        c_in = cfg.emb_width
    st["codebook"] = torch.randn(cfg.l_bins, cfg.emb_width, generator=g)
    return st


class JukeboxVQVAE:
    """Top-level (level 2) encoder + bottleneck of the Jukebox VQ-VAE."""

    def __init__(self, cfg, state, device="cuda"):
        self.cfg = cfg
        self.device = torch.device(device)
        self.w = {k: v.to(self.device, torch.float32).contiguous() for k, v in state.items()}
        self.sample_length = cfg.sample_length

    def encode_top_latents(self, x):
        """x fp32 [N,1,T] -> fp32 [N,64,T/128] (time-tiled fused res-blocks, one HBM pass per block)."""
        cfg, w = self.cfg, self.w
        h = x
        for b, down_t in enumerate(cfg.downs_t):
            for i in range(down_t):
                h = ops.jb_conv_down(h, w[f"enc.{b}.down.{i}.w"], w[f"enc.{b}.down.{i}.b"])
                for d in range(cfg.depth):
                    h = ops.jb_resblock(h, w[f"enc.{b}.res.{i}.{d}.w1"], w[f"enc.{b}.res.{i}.{d}.b1"],
                                        w[f"enc.{b}.res.{i}.{d}.w2"], w[f"enc.{b}.res.{i}.{d}.b2"],
                                        cfg.dilation_growth_rate ** d)
            h = ops.jb_conv_k3(h, w[f"enc.{b}.proj.w"], w[f"enc.{b}.proj.b"])
        return h

    def encode(self, x):
        """Reference call: vqvae.encode(FloatTensor[N,T,1]) -> zs (list, last = top level) (jukebox/main.py:61)."""
        if x.dim() != 3 or x.shape[-1] != 1:
            raise AssertionError(f"expected [N, T, 1] audio, got {tuple(x.shape)}")
        x = x.to(self.device, torch.float32).permute(0, 2, 1).contiguous()  # NTC -> NCT
        z = ops.jb_vq_encode(self.encode_top_latents(x), self.w["codebook"])
        return [None, None, z]


def random_prior_state(cfg, seed=1, device="cpu"):
    """Seeded synthetic prior weights of the upstream shapes; pass device="cuda" to create the 2.5 G parameters of
    the real 36-layer prior directly in HBM."""
    g = torch.Generator(device=device).manual_seed(seed)
    W, A, Mw = cfg.prior_width, cfg.attn_width, cfg.mlp_width

    class _R:
        @staticmethod
        def randn(*shape, generator=None):
            return torch.randn(*shape, generator=generator, device=device)

    randn = _R.randn
    st = {"x_emb": randn(cfg.l_bins, W, generator=g) * 0.02,
          "pos_emb": randn(cfg.n_ctx, W, generator=g) * 0.01,
          "x_cond": randn(cfg.n_ctx, W, generator=g) * 0.01,
          "y_cond": randn(W, generator=g) * 0.02}
    for d in range(cfg.prior_depth):
        q = f"blk.{d}."
        st[q + "ln0.g"] = 1 + 0.1 * randn(W, generator=g)
        st[q + "ln0.b"] = 0.02 * randn(W, generator=g)
        st[q + "c_attn.w"] = randn(W, 3 * A, generator=g) * (0.7 / math.sqrt(W))
        st[q + "c_attn.b"] = 0.02 * randn(3 * A, generator=g)
        st[q + "c_proj.w"] = randn(A, W, generator=g) * (0.3 / math.sqrt(A))
        st[q + "c_proj.b"] = 0.02 * randn(W, generator=g)
        st[q + "ln1.g"] = 1 + 0.1 * randn(W, generator=g)
        st[q + "ln1.b"] = 0.02 * randn(W, generator=g)
        st[q + "fc.w"] = randn(W, Mw, generator=g) * (0.7 / math.sqrt(W))
        st[q + "fc.b"] = 0.02 * randn(Mw, generator=g)
        st[q + "fc_proj.w"] = randn(Mw, W, generator=g) * (0.3 / math.sqrt(Mw))
        st[q + "fc_proj.b"] = 0.02 * randn(W, generator=g)
    return st


class PriorTransformer:
    """`top_prior.prior`: ConditionalAutoregressive2D restricted to what LLark calls (only_encode forward)."""

    def __init__(self, cfg, state, device="cuda", precision="tf32"):
        if precision not in ("tf32", "bf16", "3xtf32"):
            raise ValueError("precision must be 'tf32' (fp32 storage, single-pass TF32 tensor cores), '3xtf32' (fp32 storage, "
                             "every product composed from TF32 halves: fp32-equivalent, what the reference's fp16=False "
                             "forward computes, at three GEMM passes per linear) or 'bf16'")
        self.cfg, self.device, self.precision = cfg, torch.device(device), precision
        self.only_encode = False
        wdt = BF16 if precision == "bf16" else torch.float32
        self.w = {}
        for k, v in state.items():
            v = v.to(self.device, torch.float32)
            if k.endswith((".c_attn.w", ".c_proj.w", ".fc.w", ".fc_proj.w")):
                v = v.t().contiguous().to(wdt)  # Conv1D [n_in, n_out] -> K-major [n_out, n_in]
            self.w[k] = v.contiguous()
        if precision == "3xtf32":  # frozen weights: split once into hi (exactly TF32-representable) + lo
            for k in [k for k in self.w if k.endswith((".c_attn.w", ".c_proj.w", ".fc.w", ".fc_proj.w"))]:
                self.w[k + "_hi"], self.w[k + "_lo"] = ops.v32_split_tf32(self.w[k])

    def _lin(self, x, name, residual=None):
        if self.precision == "3xtf32":
            # x.W = x_lo.W_hi + x_hi.W_lo + x_hi.W_hi accumulated in fp32 through the GEMM's residual input (small terms
            # first); the dropped lo.lo term is ~2^-22 relative
            w_hi, w_lo = self.w[name + ".w_hi"], self.w[name + ".w_lo"]
            x_hi, x_lo = ops.v32_split_tf32(x)
            y = ops.gemm(x_lo, w_hi, residual=residual, out_dtype=torch.float32)
            y = ops.gemm(x_hi, w_lo, residual=y, out=y)
            return ops.gemm(x_hi, w_hi, residual=y, bias=self.w[name + ".b"], out=y)
        return ops.gemm(x, self.w[name + ".w"], bias=self.w[name + ".b"], residual=residual,
                        out_dtype=torch.float32)

    def forward(self, x, x_cond=None, y_cond=None, encoder_kv=None, fp16=False):
        """codes int64 [N,T] -> activations fp32 [N,T,W] (reference: jukebox/main.py:108; fp16 must be False)."""
        if not self.only_encode:
            raise NotImplementedError("only the only_encode=True activation path is part of LLark's hot path")
        if fp16 or encoder_kv is not None:
            raise NotImplementedError("LLark calls the prior with encoder_kv=None, fp16=False")
        cfg, w = self.cfg, self.w
        N, T = x.shape
        if (x_cond is None and "x_cond" not in w) or (y_cond is None and "y_cond" not in w):
            raise ValueError("x_cond / y_cond are required (top_prior.get_cond or top_prior.default_cond())")
        xc = (x_cond if x_cond is not None else w["x_cond"]).reshape(-1, cfg.prior_width)[:T].contiguous().float()
        yc = (y_cond if y_cond is not None else w["y_cond"]).reshape(-1)[: cfg.prior_width].contiguous().float()
        h = ops.jb_embed_tokens(x.to(self.device).contiguous(), w["x_emb"], yc, w["pos_emb"], xc)  # [N*T, W]
        act_dt = BF16 if self.precision == "bf16" else torch.float32
        exact = self.precision == "3xtf32"  # fp32 SIMT factored attention instead of the TF32 tensor-core kernel
        dh = cfg.attn_width // cfg.heads
        bc = T // cfg.blocks if cfg.blocks <= T else T
        for d in range(cfg.prior_depth):
            q = f"blk.{d}"
            n0 = ops.layernorm_fwd(h, w[q + ".ln0.g"], w[q + ".ln0.b"], cfg.ln_eps, out_dtype=act_dt)
            qkv = self._lin(n0, q + ".c_attn")
            a = ops.jb_sparse_attn(qkv, N, T, cfg.heads, dh, bc, d % 3, out_dtype=act_dt, exact=exact)  # cast fused (bf16)
            if a.dtype != act_dt:
                a = ops.cast_f32_to_bf16(a)
            xa = self._lin(a, q + ".c_proj", residual=h)  # x + a
            n1 = ops.layernorm_fwd(xa, w[q + ".ln1.g"], w[q + ".ln1.b"], cfg.ln_eps, out_dtype=act_dt)
            # quick-GELU (and the bf16 cast of the bf16 prior) fused into the fc GEMM's epilogue
            if exact:
                m = ops.quick_gelu(self._lin(n1, q + ".fc"))
            else:
                m = ops.gemm(n1, w[q + ".fc.w"], bias=w[q + ".fc.b"], out_dtype=act_dt, quick_gelu=True)
            h = self._lin(m, q + ".fc_proj", residual=xa)  # (x + a) + m
        return h.view(N, T, cfg.prior_width)


@dataclass
class LabelConfig:
    """prior_5b label hyper-parameters [UPSTREAM-RECALLED, hparams.py @ 08efbbc]: labels_v3, y_bins = (604 genres, 7898
    artists), t_bins = 64, one genre slot, durations 60..600 s, no lyric tokens."""
    genre_bins: int = 604
    artist_bins: int = 7898
    t_bins: int = 64
    max_bow_genre_size: int = 1
    min_duration: float = 60.0
    max_duration: float = 600.0
    sr: int = 44100
    n_tokens: int = 0


def random_label_state(cfg, lcfg=None, seed=2, device="cpu"):
    lcfg = lcfg or LabelConfig()
    g = torch.Generator(device=device).manual_seed(seed)
    W = cfg.prior_width

    def tab(n):
        return torch.randn(n, W, generator=g, device=device) * 0.01

    return {"y_emb.genre": tab(lcfg.genre_bins), "y_emb.artist": tab(lcfg.artist_bins),
            "y_emb.total_length": tab(lcfg.t_bins), "y_emb.absolute_pos": tab(lcfg.t_bins),
            "y_emb.relative_pos": tab(lcfg.t_bins)}


class _Labeller:
    """`top_prior.labeller` (upstream jukebox/data/labels.py Labeller, v3 ids): metas -> label rows
    y = [total_length, offset, sample_length, artist_id, genre ids padded with -1, (no lyric tokens for 5b)].
    `artist_ids` / `genre_ids` are the name -> id tables of upstream's ids/v3_*_ids.txt; only "unknown" -> 0 is built in
    (all the reference ever asks for, jukebox/main.py:81-89); any other name maps to 0 with a message, as upstream."""

    def __init__(self, lcfg, sample_length, artist_ids=None, genre_ids=None):
        self.lcfg, self.sample_length = lcfg, sample_length
        self.artist_ids = dict(artist_ids or {"unknown": 0})
        self.genre_ids = dict(genre_ids or {"unknown": 0})

    def _lookup(self, table, name, what):
        key = name.lower()
        if key not in table:
            print(f"Input {name} maps to {key}, which is not a known {what}. Defaulting to ({what}_id, {what}) = (0, unknown)")
        return table.get(key, 0)

    def get_label(self, artist, genre, lyrics, total_length, offset):
        n = self.lcfg.max_bow_genre_size
        genres = [self._lookup(self.genre_ids, genre, "genre")][:n]
        genres = genres + [-1] * (n - len(genres))
        y = [int(total_length), int(offset), int(self.sample_length), self._lookup(self.artist_ids, artist, "artist"),
             *genres]
        return dict(y=y, info=dict(artist=artist, genre=genre, lyrics=lyrics, full_tokens=[]))

    def get_batch_labels(self, metas, device="cpu"):
        labels = [self.get_label(**m) for m in metas]
        # label rows are index data (a few int64 per clip): they stay on the host, where get_cond bins them
        return dict(y=torch.tensor([lab["y"] for lab in labels], dtype=torch.int64), info=[lab["info"] for lab in labels])


def range_bins(pos_start, pos_end, n_time, bins, lo, hi, clamp=False):
    """Index half of upstream RangeEmbedding.forward ([UPSTREAM-RECALLED] conditioners.py): fp32 positions interpolated
    over n_time steps, normalised to [lo, hi), floored into `bins` bins. A few thousand integers per clip, computed on
    the host with the same IEEE fp32 operations the reference runs (bit-exact index work); the embedding arithmetic
    itself is llark_jb_label_cond."""
    if not ((lo <= pos_start).all() and (pos_start < hi).all()):
        raise AssertionError(f"Range is [{lo},{hi}), got {pos_start}")
    pos_start = pos_start.float()
    if pos_end is not None:
        if clamp:
            pos_end = pos_end.clamp(lo, hi)
        if not ((lo <= pos_end).all() and (pos_end <= hi).all()):
            raise AssertionError(f"Range is [{lo},{hi}), got {pos_end}")
        pos_end = pos_end.float()
    if n_time != 1:
        interpolation = torch.arange(0, n_time, dtype=torch.float).view(1, n_time) / n_time
        position = pos_start + (pos_end - pos_start) * interpolation
    else:
        position = pos_start
    normalised = (position - lo) / (hi - lo)
    return (bins * normalised).floor().long()


class JukeboxTopPrior:
    def __init__(self, cfg, state, device="cuda", precision="tf32", label_cfg=None, artist_ids=None, genre_ids=None):
        self.cfg = cfg
        self.raw_to_tokens = cfg.hop
        self.sample_length = cfg.n_ctx * cfg.hop
        self.label_cfg = label_cfg or LabelConfig()
        self.prior = PriorTransformer(cfg, state, device, precision)
        self.labeller = _Labeller(self.label_cfg, self.sample_length, artist_ids, genre_ids)
        self._cond_cache = {}

    @property
    def has_label_conditioner(self):
        return "y_emb.artist" in self.prior.w

    def get_y(self, labels, start):
        """SimplePrior.get_y: this level's sample length into column 2, the offset advanced by start * raw_to_tokens."""
        y = labels["y"].clone()
        y[:, 2] = int(self.sample_length)
        y[:, 1:2] = y[:, 1:2] + int(start * self.raw_to_tokens)
        return y

    def _cond_row(self, row):
        """(x_cond [T, W], y_cond [W]) for one label row (cached: the reference asks for the same metas for every file,
        jukebox/main.py:143, and throws 7 of its 8 computed rows away, :95-96)."""
        key = tuple(int(v) for v in row)
        hit = self._cond_cache.get(key)
        if hit is not None:
            return hit
        cfg, lc, w = self.cfg, self.label_cfg, self.prior.w
        dev = w["y_emb.artist"].device
        y = torch.tensor([key], dtype=torch.int64)
        total_length, offset, length, artist, genre = y[:, 0:1], y[:, 1:2], y[:, 2:3], y[:, 3:4], y[:, 4:]
        if genre.shape[1] != lc.max_bow_genre_size:
            raise AssertionError(f"Expected shape (N,{4 + lc.max_bow_genre_size}), got {tuple(y.shape)}")
        start, end = offset.float(), (offset + length).float()
        total = total_length.float()
        b_total = range_bins(total, None, 1, lc.t_bins, lc.min_duration * lc.sr, lc.max_duration * lc.sr)
        b_abs = range_bins(start, end, cfg.n_ctx, lc.t_bins, 0.0, lc.max_duration * lc.sr)
        b_rel = range_bins(start / total, end / total, cfg.n_ctx, lc.t_bins, 0.0, 1.0, clamp=True)
        x_cond = ops.jb_label_cond(w["y_emb.total_length"][int(b_total[0, 0])].contiguous(), w["y_emb.absolute_pos"],
                                   w["y_emb.relative_pos"], b_abs[0].to(torch.int32).to(dev),
                                   b_rel[0].to(torch.int32).to(dev))
        # start embedding: sum of the non-empty genre slots, then + artist (upstream: genre_emb + artist_emb)
        g_sum = None
        for gid in genre[0].tolist():
            if gid < 0:
                continue
            row_g = w["y_emb.genre"][gid].contiguous()
            g_sum = row_g if g_sum is None else ops.add3(g_sum, row_g)
        a_row = w["y_emb.artist"][int(artist[0, 0])].contiguous()
        y_cond = a_row if g_sum is None else ops.add3(g_sum, a_row)
        self._cond_cache = {key: (x_cond, y_cond)}  # one entry: the metas are constant in practice
        return x_cond, y_cond

    def get_cond(self, z_conds, y):
        """SimplePrior.get_cond for the top level -> (x_cond [n,T,W], y_cond [n,1,W], prime=None): with the label tables
        in the state (a real checkpoint, or random_label_state) the conditioning is computed from the label rows `y`;
        a state that carries ready-made "x_cond" / "y_cond" constants (round-1 synthetic states) uses those."""
        w = self.prior.w
        if not self.has_label_conditioner:
            n = y.shape[0] if torch.is_tensor(y) else getattr(y, "n", 1)
            return w["x_cond"][None].expand(n, -1, -1), w["y_cond"][None, None].expand(n, -1, -1), None
        rows = [tuple(r) for r in y.tolist()]
        if all(r == rows[0] for r in rows):
            x, yc = self._cond_row(rows[0])
            return x[None].expand(len(rows), -1, -1), yc[None, None].expand(len(rows), -1, -1), None
        pairs = [self._cond_row(r) for r in rows]
        return torch.stack([p[0] for p in pairs]), torch.stack([p[1] for p in pairs])[:, None], None

    def default_cond(self):
        """Conditioning of the reference's fixed metas (jukebox/main.py:71-98) as (x_cond [T,W], y_cond [W])."""
        if not self.has_label_conditioner:
            return self.prior.w["x_cond"], self.prior.w["y_cond"]
        total = (int(62 * self.label_cfg.sr) // self.raw_to_tokens) * self.raw_to_tokens
        meta = dict(artist="unknown", genre="unknown", total_length=total, offset=0, lyrics="lyrics go here!!!")
        y = self.get_y(self.labeller.get_batch_labels([meta]), 0)
        return self._cond_row(tuple(y[0].tolist()))


# ------------------------------------------------------------------------------------------------ upstream checkpoints
def vqvae_key_map(cfg, level=2):
    """[UPSTREAM-RECALLED] vqvae.pth.tar names of the level-2 encoder and codebook -> this module's names
    (Encoder.level_blocks[b].model = Sequential(Sequential(Conv1d k4 s2, Resnet1D) x down_t, Conv1d k3);
    ResConv1DBlock.model = Sequential(ReLU, Conv1d k3 dilated, ReLU, Conv1d k1))."""
    m = {}
    for b, down_t in enumerate(cfg.downs_t[: level + 1]):
        base = f"encoders.{level}.level_blocks.{b}.model."
        for i in range(down_t):
            for s, t in (("weight", "w"), ("bias", "b")):
                m[f"{base}{i}.0.{s}"] = f"enc.{b}.down.{i}.{t}"
                for d in range(cfg.depth):
                    m[f"{base}{i}.1.model.{d}.model.1.{s}"] = f"enc.{b}.res.{i}.{d}.{t}1"
                    m[f"{base}{i}.1.model.{d}.model.3.{s}"] = f"enc.{b}.res.{i}.{d}.{t}2"
        for s, t in (("weight", "w"), ("bias", "b")):
            m[f"{base}{down_t}.{s}"] = f"enc.{b}.proj.{t}"
    m[f"bottleneck.level_blocks.{level}.k"] = "codebook"
    return m


def prior_key_map(cfg):
    """[UPSTREAM-RECALLED] prior_level_2.pth.tar names -> this module's names, first cfg.prior_depth blocks only."""
    m = {"prior.x_emb.weight": "x_emb", "prior.pos_emb.pos_emb": "pos_emb",
         "y_emb.bow_genre_emb.emb.weight": "y_emb.genre", "y_emb.artist_emb.emb.weight": "y_emb.artist",
         "y_emb.total_length_emb.emb.weight": "y_emb.total_length",
         "y_emb.absolute_pos_emb.emb.weight": "y_emb.absolute_pos",
         "y_emb.relative_pos_emb.emb.weight": "y_emb.relative_pos"}
    for d in range(cfg.prior_depth):
        u, q = f"prior.transformer._attn_mods.{d}.", f"blk.{d}."
        m.update({u + "ln_0.weight": q + "ln0.g", u + "ln_0.bias": q + "ln0.b",
                  u + "attn.c_attn.w": q + "c_attn.w", u + "attn.c_attn.b": q + "c_attn.b",
                  u + "attn.c_proj.w": q + "c_proj.w", u + "attn.c_proj.b": q + "c_proj.b",
                  u + "ln_1.weight": q + "ln1.g", u + "ln_1.bias": q + "ln1.b",
                  u + "mlp.c_fc.w": q + "fc.w", u + "mlp.c_fc.b": q + "fc.b",
                  u + "mlp.c_proj.w": q + "fc_proj.w", u + "mlp.c_proj.b": q + "fc_proj.b"})
    return m


def load_upstream_state(path, key_map):
    """restore_model of upstream make_models.py as patched by the reference (jukebox/make_models.py.patch:7-8):
    torch.load(path)["model"], a leading "module." stripped from every name, then a NON-STRICT load -- names the
    36-block model does not have (blocks 36..71 of the 72-block prior, the other VQ-VAE levels, decoders, x_out, ...) are
    ignored; names the model has but the file lacks are reported. -> (state dict in this module's names, missing names)."""
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    model = ckpt["model"] if isinstance(ckpt, dict) and "model" in ckpt else ckpt
    model = {(k[7:] if k[:7] == "module." else k): v for k, v in model.items()}
    state = {ours: model[theirs] for theirs, ours in key_map.items() if theirs in model}
    missing = [theirs for theirs in key_map if theirs not in model]
    return state, missing
