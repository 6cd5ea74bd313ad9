"""Execution plans of the hot path, written against *duck-typed* modules.

Everything here reads a module only through the attributes the reference's classes define
(``in_stem / down_layers / pre_enc_layers / vq_layers / shortcut_layers`` on an encoder
(vq_ae/model.py:129-187), ``out_stem / up_layers / post_enc_layers`` on a decoder (:220-272),
``embed / proj_in / proj_out / commitment_cost`` on a quantiser (layers/vq.py:17-42,168-187),
``bias1a .. scale / branch_conv1..3 / skip_conv`` on a Fixup block (layers/conv_block.py:136-194)).
So the same functions serve this package's mirrored classes (model.py, layers/*.py) and instances
of the UNMODIFIED reference classes that ``vqae_b200.accelerate`` has bound to the B200 path.

Per-module state (packed weights, selected precision, near-tie counter) lives in
``module.__dict__['_b200']`` -- no parameters, no buffers, ``state_dict`` is untouched.
"""
from __future__ import annotations

import types
from math import prod
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor, nn

from . import engine as E
from . import mbconv as M

REDUCED = "fp16"        # what an active torch.autocast('cuda') selects


def state(module) -> types.SimpleNamespace:
    st = module.__dict__.get("_b200")
    if st is None:
        st = types.SimpleNamespace(precision=None, plan=None, packed=None, packed_key=None,
                                   last_near_ties=None)
        module.__dict__["_b200"] = st
    return st


def resolve_precision(module) -> str:
    """Arithmetic of a plan run: an explicit ``set_precision`` wins; otherwise an active
    ``torch.autocast('cuda')`` (the reference's extraction loop, extract_embeddings.py:124-125,
    and eval.py:44) selects the reduced-precision tensor-core path, and plain calls run fp32 --
    the same rule the reference's modules follow under PyTorch."""
    p = getattr(module, "precision", None) or state(module).precision
    if p is not None and p != "auto":
        return p
    return REDUCED if torch.is_autocast_enabled("cuda") else "fp32"


def is_fixup_block(m) -> bool:
    return all(hasattr(m, a) for a in ("branch_conv1", "branch_conv2", "branch_conv3", "bias1a",
                                       "scale"))


def flat_blocks(module: nn.Module) -> List[nn.Module]:
    """All PreActFixupResBlocks below ``module`` in execution order."""
    return [m for m in module.modules() if is_fixup_block(m)]


class Plan:
    """Packed blocks of one Sequential chain, re-packed when any parameter changes."""

    def __init__(self):
        self.key = None
        self.packed: List[E.PackedFixup] = []
        self.chains: dict = {}

    def get(self, blocks: Sequence[nn.Module]) -> List[E.PackedFixup]:
        key = tuple(E.block_version(b) for b in blocks)
        if key != self.key:
            for b in blocks:
                E.check_block_supported(b)
            self.packed = E.pack_blocks(blocks)
            self.chains = {}
            self.key = key
        return self.packed

    def run(self, blocks: Sequence[nn.Module], h: Tensor, precision: str = "fp32") -> Tensor:
        return E.run_blocks_nhwc(self.get(blocks), h, precision, self.chains)

    def run_from_input(self, stem, blocks: Sequence[nn.Module], x: Tensor, mean, std,
                       precision: str, half_stream: bool) -> Tensor:
        """in_stem + the chain; the stem and the leading blocks fuse into one launch where a kernel
        is built for them (E.encoder_front)."""
        packed = self.get(blocks)
        h, used = E.encoder_front(x, stem.weight, stem.bias, mean, std, packed, precision, half_stream)
        return E.run_blocks_nhwc(packed[used:], h, precision, self.chains)


def _plan(module) -> Plan:
    st = state(module)
    if st.plan is None:
        st.plan = Plan()
    return st.plan


# ---- quantiser (layers/vq.py) ------------------------------------------------------------------
def packed_quantizer(vq) -> E.PackedQuantizer:
    proj_in, proj_out = getattr(vq, "proj_in", None), getattr(vq, "proj_out", None)
    tensors = [vq.embed] + ([proj_in.weight, proj_in.bias, proj_out.weight, proj_out.bias]
                            if proj_in is not None else [])
    key = tuple((t.data_ptr(), t._version) for t in tensors) + (float(vq.commitment_cost),)
    st = state(vq)
    if st.packed is None or key != st.packed_key:
        st.packed = E.PackedQuantizer(vq.embed, vq.commitment_cost, proj_in, proj_out)
        st.packed_key = key
    return st.packed


def check_quantizer_input(vq, inputs: Tensor) -> None:
    """The argument checks of EMAVectorQuantizer.forward (vq.py:97-104) + what is built."""
    channels = vq.proj_in.in_channels if getattr(vq, "proj_in", None) is not None \
        else vq.embedding_dim
    ndim = inputs.dim()
    assert ndim >= 3                                                    # vq.py:98
    if inputs.shape[1] != channels:                                     # vq.py:100-104
        raise NotImplementedError(
            'VQ dim != channel dim not supported;'
            f' found channel dim of {inputs.shape[1]}, expected {channels}')
    if ndim != 4:
        # the reference passes p = inputs.dim() to cdist (vq.py:121-129); only p = 4 is built
        raise NotImplementedError(f"only 4-D inputs (L4 distance) are supported, got {ndim}-D")
    E.require_cuda(inputs, type(vq).__name__ + ".forward")


def quantizer_forward(vq, inputs: Tensor, want_z: bool = False):
    """(quantized, indices int64, loss 0-dim[, z]) of vq.py:96-154 / 185-192 in eval mode."""
    check_quantizer_input(vq, inputs)
    pq = packed_quantizer(vq)
    b = inputs.shape[0]
    s = prod(inputs.shape[2:])
    cl = E.is_channels_last(inputs)
    x = inputs if inputs.dtype == torch.float32 else inputs.float()
    x = x.permute(0, 2, 3, 1).contiguous() if cl else x.contiguous()
    out, idx, loss, ties, z = E.quantize(pq, x, cl, cl, b, s, want_out=True, want_z=want_z)
    state(vq).last_near_ties = ties
    sp = tuple(inputs.shape[2:])
    quantized = out.view(b, *sp, pq.c).permute(0, 3, 1, 2) if cl else out.view(b, pq.c, *sp)
    if inputs.dtype != torch.float32:
        quantized = quantized.to(inputs.dtype)
    return quantized, idx.view(b, *sp), loss, z


def decode_codes(vq, embed_idx: Tensor, channels_last: bool = False) -> Tensor:
    """proj_out(embed_code(idx)) for stored code maps: [B,H,W] int -> [B,C,H,W]."""
    E.require_cuda(embed_idx, "decode_codes")
    pq = packed_quantizer(vq)
    b, h, w = embed_idx.shape
    out = E.embed_codes(pq, embed_idx.reshape(-1), channels_last, b, h * w)
    if channels_last:
        return out.view(b, h, w, pq.c).permute(0, 3, 1, 2)
    return out.view(b, pq.c, h, w)


def _world() -> int:
    import torch.distributed as dist
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def quantizer_forward_training(vq, inputs: Tensor):
    """EMAVectorQuantizer.forward in TRAINING mode (vq.py:96-154; scope row f-4): on the first pass the
    codebook is re-centred on the batch statistics (``_init_ema``, vq.py:76-94), the nearest codes are
    searched with the codebook as it stands, and the EMA buffers are updated from this batch
    (``_update_ema``, vq.py:47-74) -- ``quantized`` and ``loss`` belong to the codebook BEFORE the update,
    like the reference.  Under ``torch.distributed`` the batch statistics are all-reduced exactly where the
    reference does it (vq.py:56-58, 81-88), as ONE collective per step instead of two.
    The search and the update run under no_grad in the reference too (vq.py:106); what is NOT built is the
    autograd graph of the straight-through estimator and of the loss: tensors come back detached."""
    import torch.distributed as dist
    check_quantizer_input(vq, inputs)
    b, s = inputs.shape[0], prod(inputs.shape[2:])
    cl = E.is_channels_last(inputs)
    x = inputs.detach()
    x = x if x.dtype == torch.float32 else x.float()
    x = x.permute(0, 2, 3, 1).contiguous() if cl else x.contiguous()
    k, d = vq.embed.shape
    world = _world()
    if bool(vq.first_pass):                                             # vq.py:118-119
        _, _, _, _, z = E.quantize(packed_quantizer(vq), x, cl, cl, b, s, want_out=False, want_z=True)
        mean, std = E.column_stats(z)
        if world > 1:                                                   # vq.py:81-88
            ms = torch.cat([mean, std])
            dist.all_reduce(ms)
            ms /= world
            mean, std = ms[:d].contiguous(), ms[d:].contiguous()
        E.ema_init(vq.embed, vq.embed_avg, vq.cluster_size, mean, std, (b * s * world) / k)
        vq.first_pass.mul_(0)                                           # vq.py:94
        state(vq).packed = None                                         # the codebook moved
    pq = packed_quantizer(vq)
    out, idx, loss, ties, z = E.quantize(pq, x, cl, cl, b, s, want_out=True, want_z=True)
    acc = E.ema_accumulate(z, idx, k)                                   # vq.py:49-54
    if world > 1:
        dist.all_reduce(acc)                                            # vq.py:56-58
    E.ema_update(vq.embed, vq.embed_avg, vq.cluster_size, acc, vq.decay, vq.laplace_alpha)
    state(vq).packed = None
    state(vq).last_near_ties = ties
    sp = tuple(inputs.shape[2:])
    quantized = out.view(b, *sp, pq.c).permute(0, 3, 1, 2) if cl else out.view(b, pq.c, *sp)
    if inputs.dtype != torch.float32:
        quantized = quantized.to(inputs.dtype)
    return quantized, idx.view(b, *sp), loss


# ---- encoder / decoder (model.py) --------------------------------------------------------------
def _single_level(levels: int, shortcuts) -> bool:
    return levels == 1 and all(sc is None for sc in shortcuts)


def _fused_plan_applies(module, levels: int, shortcuts) -> bool:
    """The one-plan fast path: one VQ level, no shortcut blocks, Fixup blocks only."""
    return _single_level(levels, shortcuts) and not any(M.is_mbconv(m) for m in module.modules())


def _check_single_level(levels: int, shortcuts, who: str) -> None:
    if not _single_level(levels, shortcuts):
        raise NotImplementedError(
            f"{who} is the single-level entry point (conf/model/vq_ae.yaml); multi-level hierarchies "
            "go through forward() (plan.encoder_forward_levels / decoder_forward_levels)")


def compute_blocks(module: nn.Module) -> List[nn.Module]:
    """The residual blocks below ``module`` in execution order: PreActFixupResBlocks and MBConvs."""
    out: List[nn.Module] = []

    def walk(m):
        if is_fixup_block(m) or M.is_mbconv(m):
            out.append(m)
            return
        for child in m.children():
            walk(child)

    walk(module)
    return out


def _run_module(module, h: Tensor, precision: str) -> Tensor:
    """A shortcut / pyramid / trunk module on an NHWC tensor: every container the reference builds these
    from (DownBlock, UpBlock, EnvelopBlock, nn.Sequential, a bare block) is a chain of residual blocks in
    ``modules()`` order.  Runs of Fixup blocks execute as packed plans (fused / resident kernels where
    they apply); MBConv blocks (fp32 kernels in every precision mode) one by one."""
    blocks = compute_blocks(module)
    leaves = [m for m in module.modules() if not list(m.children())]
    inside = {id(l) for blk in blocks for l in blk.modules()}
    foreign = [type(l).__name__ for l in leaves if id(l) not in inside
               and not (isinstance(l, (nn.Sequential, nn.ModuleList)) and len(l) == 0)]
    if foreign:
        raise NotImplementedError(
            f"{type(module).__name__}: only chains of PreActFixupResBlocks / MBConvs have B200 kernels "
            f"(found {sorted(set(foreign))})")
    st = state(module)
    if not hasattr(st, "runs"):
        st.runs = {}
    i = 0
    while i < len(blocks):
        if M.is_mbconv(blocks[i]):
            h = M.forward_nhwc(M.packed(blocks[i]), h)
            i += 1
            continue
        j = i
        while j < len(blocks) and is_fixup_block(blocks[j]):
            j += 1
        plan = st.runs.get(i)
        if plan is None:
            plan = st.runs[i] = Plan()
        h = plan.run(blocks[i:j], h, precision)
        i = j
    return h


def _level_out(out: Tensor, b: int, hh: int, ww: int, c: int, cl: bool) -> Tensor:
    return out.view(b, hh, ww, c).permute(0, 3, 1, 2) if cl else out.view(b, c, hh, ww)


def encoder_forward_levels(enc, x: Tensor, mean=None, std=None):
    """Encoder.forward for any number of VQ levels (model.py:189-217): the pyramid of DownBlocks, then
    from the LOWEST resolution up: ``enc(pre_enc(down + shortcut(aux[0])))`` with ``aux`` the previous
    (lower) level's quantiser output.  Returns ((enc..), (idx..), (loss..)), low-res first."""
    if enc.training:
        raise RuntimeError("Encoder: training-mode forward is outside the B200 inference "
                           "path; call .eval()")
    E.require_cuda(x, "Encoder.forward")
    precision = resolve_precision(enc)
    cl = x.dtype == torch.uint8 or E.is_channels_last(x)
    downs: List[Tensor] = []
    first = compute_blocks(enc.down_layers[0])
    if first and all(is_fixup_block(b) for b in first):
        h = _plan(enc.down_layers[0]).run_from_input(enc.in_stem, first, x, mean, std, precision, False)
    else:
        h = _run_module(enc.down_layers[0],
                        E.stem_in(x, enc.in_stem.weight, enc.in_stem.bias, mean, std, precision=precision),
                        precision)
    downs.append(h)
    for down_layer in list(enc.down_layers)[1:]:
        h = _run_module(down_layer, h, precision)
        downs.append(h)
    outs = []
    aux: Optional[Tensor] = None                        # NHWC quantiser output of the previous level
    for down, pre_enc, vq, shortcut in zip(reversed(downs), enc.pre_enc_layers, enc.vq_layers,
                                           enc.shortcut_layers):
        if shortcut is not None:
            down = E.add_nhwc(down, _run_module(shortcut, aux, precision))      # model.py:208
        h = _run_module(pre_enc, down, precision)
        pq = packed_quantizer(vq)
        b, hh, ww, c = h.shape
        if c != pq.c:
            raise NotImplementedError(
                'VQ dim != channel dim not supported;'
                f' found channel dim of {c}, expected {pq.c}')
        out, idx, loss, ties, _ = E.quantize(pq, h.float(), True, True, b, hh * ww, want_out=True)
        state(vq).last_near_ties = ties
        aux = out.view(b, hh, ww, c)
        q = aux.permute(0, 3, 1, 2)
        outs.append((q if cl else q.contiguous(), idx.view(b, hh, ww), loss))
    return tuple(zip(*outs))


def decoder_forward_levels(dec, xs: Sequence[Tensor]) -> Tensor:
    """Decoder.forward for any number of levels (model.py:274-291), x low-res first:
    ``prev_up = up(prev_up + post_enc(shortcut(aux) + enc))`` with ``aux`` the previous level's ``enc``."""
    if dec.training:
        raise RuntimeError("Decoder: training-mode forward is outside the B200 inference "
                           "path; call .eval()")
    if len(xs) != len(dec.up_layers):
        raise ValueError(f"Decoder: {len(xs)} encodings for {len(dec.up_layers)} levels")
    precision = resolve_precision(dec)
    prev_up: Optional[Tensor] = None
    aux: Optional[Tensor] = None
    cl = False
    for enc_t, shortcut, post_enc, up in zip(xs, dec.shortcut_layers, dec.post_enc_layers,
                                             dec.up_layers):
        E.require_cuda(enc_t, "Decoder.forward")
        e, cl = E.to_nhwc(enc_t)
        h = e if shortcut is None else E.add_nhwc(_run_module(shortcut, aux, precision), e)
        aux = e                                                          # model.py:286
        h = _run_module(post_enc, h, precision)
        if prev_up is not None:
            h = E.add_nhwc(prev_up, h)                                   # model.py:283
        prev_up = _run_module(up, h, precision)
    return E.stem_out(prev_up, dec.out_stem.weight, dec.out_stem.bias, cl, precision)


def encoder_encode(enc, x: Tensor, mean=None, std=None, want_quantized: bool = True,
                   want_latents: bool = False):
    """Run the encoder plan.  x: float [B,3,H,W] (NCHW or channels_last) or uint8 [B,H,W,3]
    (normalised on the fly, a-N fused into the stem).  Returns
    (enc or None, indices int64 [B,h,w], loss 0-dim, near_ties 0-dim int32, z or None)."""
    _check_single_level(len(enc.vq_layers), enc.shortcut_layers, "Encoder.encode")
    if enc.training:
        raise RuntimeError("Encoder: training-mode forward is outside the B200 inference "
                           "path; call .eval()")
    E.require_cuda(x, "Encoder.forward")
    precision = resolve_precision(enc)
    cl = x.dtype == torch.uint8 or E.is_channels_last(x)
    half = precision == "fp16" and E.STREAM_F16
    # one plan over stem + pyramid + trunk: the 'same' blocks that close the last DownBlock and the
    # trunk are one run of equal-width blocks, i.e. ONE image-resident launch; in_stem and the first
    # two blocks are one launch in the reduced-precision mode
    h = _plan(enc).run_from_input(enc.in_stem,
                                  flat_blocks(enc.down_layers) + flat_blocks(enc.pre_enc_layers),
                                  x, mean, std, precision, half)
    vq = enc.vq_layers[0]
    pq = packed_quantizer(vq)
    b, hh, ww, c = h.shape
    if c != pq.c:
        raise NotImplementedError(
            'VQ dim != channel dim not supported;'
            f' found channel dim of {c}, expected {pq.c}')
    if h.dtype != torch.float32 and not (E.quantize_io_supported(pq, h.dtype) and cl):
        h = h.float()
    out, idx, loss, ties, z = E.quantize(pq, h, True, cl, b, hh * ww,
                                         want_out=want_quantized, want_z=want_latents)
    if out is not None and out.dtype != torch.float32:
        out = out.float()                    # the module API returns fp32 like the reference
    state(vq).last_near_ties = ties
    enc_t = None
    if out is not None:
        enc_t = (out.view(b, hh, ww, c).permute(0, 3, 1, 2) if cl else out.view(b, c, hh, ww))
    return enc_t, idx.view(b, hh, ww), loss, ties, (z.view(b, hh, ww, -1) if z is not None
                                                     else None)


def decoder_forward(dec, xs: Sequence[Tensor]) -> Tensor:
    if not _fused_plan_applies(dec, len(dec.up_layers), dec.shortcut_layers):
        return decoder_forward_levels(dec, xs)
    if len(xs) != 1:
        raise ValueError(f"Decoder: {len(xs)} encodings for 1 level")
    if dec.training:
        raise RuntimeError("Decoder: training-mode forward is outside the B200 inference "
                           "path; call .eval()")
    enc = xs[0]
    E.require_cuda(enc, "Decoder.forward")
    precision = resolve_precision(dec)
    h, cl = E.to_nhwc(enc)
    h = _plan(dec).run(flat_blocks(dec.post_enc_layers) + flat_blocks(dec.up_layers), h, precision)
    return E.stem_out(h, dec.out_stem.weight, dec.out_stem.bias, cl, precision)


def block_forward(block, inp: Tensor) -> Tensor:
    """One PreActFixupResBlock on an NCHW / channels_last tensor (conv_block.py:196-216)."""
    if block.training:
        raise RuntimeError("PreActFixupResBlock: training-mode forward is outside the B200 "
                           "inference path; call .eval()")
    E.require_cuda(inp, "PreActFixupResBlock.forward")
    st = state(block)
    key = E.block_version(block)
    if st.packed is None or key != st.packed_key:
        E.check_block_supported(block)
        st.packed = E.pack_blocks([block])[0]
        st.packed_key = key
    x, cl = E.to_nhwc(inp)
    return E.from_nhwc(E.fixup_forward_nhwc(st.packed, x, precision=resolve_precision(block)), cl)


# ---- binding onto instantiated reference modules -------------------------------------------------
def _bind(module, fast):
    """Route eval-mode CUDA forwards of ``module`` to ``fast``; training mode and CPU tensors keep
    the module's own (reference) forward."""
    if "_b200_orig_forward" in module.__dict__:
        return
    orig = module.forward

    def forward(self, *args, **kwargs):
        first = args[0] if args else None
        if isinstance(first, (list, tuple)) and first:
            first = first[0]
        if self.training or not (isinstance(first, Tensor) and first.is_cuda):
            return orig(*args, **kwargs)
        return fast(self, *args, **kwargs)

    module.__dict__["_b200_orig_forward"] = orig
    module.forward = types.MethodType(forward, module)


def encoder_forward(enc, x):
    """Encoder.forward (model.py:189-217): the single-level fast plan, or the level loop."""
    if not _fused_plan_applies(enc, len(enc.vq_layers), enc.shortcut_layers):
        return encoder_forward_levels(enc, x)
    e, idx, loss, _, _ = encoder_encode(enc, x)
    return (e,), (idx,), (loss,)


_enc_fast = encoder_forward


def _vq_fast(vq, inputs):
    q, idx, loss, _ = quantizer_forward(vq, inputs)
    return q, idx, loss


def accelerate(model: nn.Module, precision: Optional[str] = None) -> nn.Module:
    """Bind an instantiated model built from the reference's OWN classes (``vq_ae.model.VQAE`` /
    ``Encoder`` / ``Decoder``, the quantisers of ``vq_ae.layers.vq``, ``PreActFixupResBlock``) to the
    B200 path, in place: eval-mode forwards on CUDA tensors run the packed plans of this module,
    everything else (training, CPU tensors, ``state_dict``, checkpoints, attributes) is the
    reference's code, untouched.  ``precision``: "fp32", "fp16", ... or None = follow
    ``torch.autocast`` like the reference does."""
    if precision is not None and precision not in E.PRECISIONS:
        raise ValueError(f"precision must be one of {E.PRECISIONS}")
    n = 0
    for m in model.modules():
        if hasattr(m, "in_stem") and hasattr(m, "vq_layers") and hasattr(m, "pre_enc_layers"):
            _bind(m, _enc_fast)
            m.encode = types.MethodType(encoder_encode, m)
        elif hasattr(m, "out_stem") and hasattr(m, "up_layers") and hasattr(m, "post_enc_layers"):
            _bind(m, decoder_forward)
        elif hasattr(m, "embed") and hasattr(m, "commitment_cost") and hasattr(m, "embed_code"):
            _bind(m, _vq_fast)
        elif is_fixup_block(m):
            _bind(m, block_forward)
        elif M.is_mbconv(m):
            _bind(m, M.block_forward)
        else:
            continue
        state(m).precision = precision
        n += 1
    if n == 0:
        raise TypeError("accelerate(): no Encoder / Decoder / quantiser / Fixup block found below "
                        f"{type(model).__name__}")
    return model
