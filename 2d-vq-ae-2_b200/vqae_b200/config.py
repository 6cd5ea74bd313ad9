"""Hand-composed model configuration, equal to the resolved Hydra tree of the reference.

The reference builds its model from ``conf/model/vq_ae.yaml`` through Hydra; Hydra is
not a dependency of this package, so the resolved tree is composed here as plain dicts
with the reference's own ``_target_`` strings (they resolve to this package once
``vqae_b200.install_as_vq_ae()`` has run, or to the reference itself under
``oracle/ref_shim.py``).

Sources (all relative to the reference checkout):
  conf/model/vq_ae.yaml:16-44, conf/model/encoder/default.yaml:1-16,
  conf/model/decoder/default.yaml:1-12,
  conf/model/layers/conv_block/pre_activation_fixup.yaml:1-74,
  conf/model/layers/conv_block/conv_layer/{conv2d,proj2d,same2d,down2d,out2d,up2dresize}.yaml,
  conf/model/layers/conv_block/{down_block,up_block}.yaml,
  conf/model/layers/vq/projected_ema_vq_2d.yaml:1-7,
  conf/model/optional_overrides/pre_activation_fixup/n_layers.yaml:3.
"""
from __future__ import annotations

from copy import deepcopy
from typing import Any, Dict, Optional


def _conv2d(**over) -> Dict[str, Any]:
    base = dict(
        _target_="torch.nn.Conv2d", in_channels=None, out_channels=None, kernel_size=None,
        stride=1, padding=0, dilation=1, groups=1, bias=True, padding_mode="zeros",
    )
    base.update(over)
    return base


def proj2d(**over):
    return _conv2d(kernel_size=1, stride=1, padding=0, **over)


def same2d(**over):
    return _conv2d(kernel_size=3, stride=1, padding=1, **over)


def down2d(**over):
    return _conv2d(kernel_size=2, stride=2, padding=0, **over)


def out2d(**over):
    return _conv2d(kernel_size=3, stride=1, padding=1, **over)


def up2dresize(**over):
    conf = proj2d(**over)
    conf["_target_"] = "vq_ae.layers.conv.ResizeConv2D"
    return conf


def pre_activation_fixup(n_layers: Optional[int]) -> Dict[str, Any]:
    nb = dict(bias=False)
    circ = dict(bias=False, padding_mode="circular")
    return dict(
        _target_="vq_ae.layers.conv_block.PreActFixupResBlock",
        _recursive_=False,
        in_channels=None, out_channels=None, mode=None, n_layers=n_layers,
        bottleneck_divisor=1,
        activation=dict(_target_="torch.nn.ELU", alpha=1.0),
        conv_conf=dict(
            down=dict(branch_conv1=proj2d(**nb), branch_conv2=down2d(**circ),
                      branch_conv3=proj2d(**nb), skip_conv=down2d(**circ)),
            up=dict(branch_conv1=proj2d(**nb), branch_conv2=up2dresize(**nb),
                    branch_conv3=proj2d(**nb), skip_conv=up2dresize(**nb)),
            same=dict(branch_conv1=proj2d(**nb), branch_conv2=same2d(**circ),
                      branch_conv3=proj2d(**nb), skip_conv=proj2d(**nb)),
            out=dict(branch_conv1=proj2d(**nb), branch_conv2=out2d(**circ),
                     branch_conv3=proj2d(**nb), skip_conv=out2d(**nb)),
        ),
    )


def projected_ema_vq_2d(embedding_dim: int, num_embeddings: int = 256,
                        projection_dim: int = 8) -> Dict[str, Any]:
    return dict(
        _target_="vq_ae.layers.vq.ProjectedEMAVectorQuantizer2d",
        num_embeddings=num_embeddings, embedding_dim=embedding_dim, commitment_cost=1,
        decay=0.99, laplace_alpha=1e-5, projection_dim=projection_dim,
    )


def compose_vqae_conf(
    n_down: int = 4,
    n_pre_layers: int = 1,
    n_post_layers: int = 4,
    n_enc_layers_trunk: int = 50,
    stem_channels: int = 8,
    in_channels: int = 3,
    num_embeddings: int = 256,
    projection_dim: int = 8,
    fixup_n_layers: Optional[int] = -1,
) -> Dict[str, Any]:
    """Resolved ``model`` config.  ``n_down=4`` is the as-shipped 512^2 model
    (conf/model/vq_ae.yaml:26); ``n_down=3`` is the README's 256^2 -> 32x32 model.

    ``fixup_n_layers=-1`` applies the reference's formula
    (optional_overrides/pre_activation_fixup/n_layers.yaml:3); ``None`` skips the Fixup
    initialisation (default torch init), an int overrides it.
    """
    if fixup_n_layers == -1:
        fixup_n_layers = (n_down * n_pre_layers * n_post_layers + n_enc_layers_trunk
                          + n_enc_layers_trunk + n_down * n_pre_layers * n_post_layers)
    fixup = pre_activation_fixup(fixup_n_layers)
    c_lat = stem_channels * 2 ** n_down

    encoder_conf = dict(
        _target_="vq_ae.model.Encoder", _recursive_=False,
        stem_conf=same2d(in_channels=in_channels, out_channels=stem_channels),
        down_block_conf=dict(
            _target_="vq_ae.layers.conv_block.DownBlock", _recursive_=False,
            in_channels=None, n_down=n_down, n_pre_layers=n_pre_layers,
            n_post_layers=n_post_layers, conv_conf=deepcopy(fixup),
        ),
        conv_block_conf=deepcopy(fixup),
        shortcut_block_conf=None,
        vq_conf={
            "_target_": "utils.conf_helpers.instantiate_dictified_listconf",
            "_recursive_": False,
            "0": projected_ema_vq_2d(c_lat, num_embeddings, projection_dim),
        },
        n_pre_enc_layers=n_enc_layers_trunk,
    )
    decoder_conf = dict(
        _target_="vq_ae.model.Decoder", _recursive_=False,
        n_enc_layers=1,
        stem_conf=same2d(in_channels=stem_channels, out_channels=in_channels),
        up_block_conf=dict(
            _target_="vq_ae.layers.conv_block.UpBlock", _recursive_=False,
            out_channels=None, n_up=n_down, n_pre_layers=n_pre_layers,
            n_post_layers=n_post_layers, conv_conf=deepcopy(fixup),
        ),
        conv_block_conf=deepcopy(fixup),
        shortcut_block_conf=None,
        n_post_enc_layers=n_enc_layers_trunk,
    )
    return dict(
        _target_="vq_ae.model.VQAE", _recursive_=False,
        optim_conf=dict(_target_="torch.optim.AdamW", lr=1e-4),
        loss_f_conf=dict(_target_="torch.nn.modules.loss.HuberLoss", reduction="mean", delta=1.0),
        encoder_conf=encoder_conf,
        decoder_conf=decoder_conf,
    )


def compose_multilevel_conf(
    level_downs=(2, 1),
    n_pre_enc_layers=(2, 3),
    n_pre_layers: int = 1,
    n_post_layers: int = 1,
    stem_channels: int = 8,
    in_channels: int = 3,
    num_embeddings: int = 256,
    projection_dim: int = 8,
    fixup_n_layers: Optional[int] = 12,
    shortcut_mode: str = "up",
) -> Dict[str, Any]:
    """A MULTI-LEVEL hierarchy in the reference's own configuration language (scope row f-4): one
    DownBlock / VQ layer / pre-enc trunk per level and a Fixup shortcut block from each lower level to
    the one above it (model.py:144-187, 203-215; conf/model/encoder/default.yaml:5 is where a shortcut
    conf would be plugged in -- the shipped tree sets it to null).

    ``level_downs[i]`` = n_down of level i's DownBlock, listed high-res first like ``vq_conf`` '0', '1', ...
    The encoder is valid in the reference for any ``level_downs``.  The reference's Decoder hands every
    shortcut block the channel count of the level it comes FROM (model.py:252-256 after the
    prepended-None shift), so a decoder only exists where consecutive levels have equal widths, i.e.
    ``level_downs[i > 0] == 0`` with ``shortcut_mode='same'`` -- ``decoder_conf`` is None otherwise."""
    fixup = pre_activation_fixup(fixup_n_layers)
    n_levels = len(level_downs)
    widths, c = [], stem_channels
    for nd in level_downs:
        c *= 2 ** nd
        widths.append(c)

    def down(nd):
        return dict(_target_="vq_ae.layers.conv_block.DownBlock", _recursive_=False, in_channels=None,
                    n_down=nd, n_pre_layers=n_pre_layers, n_post_layers=n_post_layers,
                    conv_conf=deepcopy(fixup))

    def up(nu):
        return dict(_target_="vq_ae.layers.conv_block.UpBlock", _recursive_=False, out_channels=None,
                    n_up=nu, n_pre_layers=n_pre_layers, n_post_layers=n_post_layers,
                    conv_conf=deepcopy(fixup))

    # encoder shortcut i: from level i+1 (lower) to level i; instantiated with in_channels = widths[i+1]
    enc_shortcuts = [{**deepcopy(fixup), "mode": shortcut_mode, "out_channels": widths[i]}
                     for i in range(n_levels - 1)]
    vq_conf: Dict[str, Any] = {"_target_": "utils.conf_helpers.instantiate_dictified_listconf",
                               "_recursive_": False}
    for i, w in enumerate(widths):
        vq_conf[str(i)] = projected_ema_vq_2d(w, num_embeddings, projection_dim)
    encoder_conf = dict(
        _target_="vq_ae.model.Encoder", _recursive_=False,
        stem_conf=same2d(in_channels=in_channels, out_channels=stem_channels),
        down_block_conf=[down(nd) for nd in level_downs],
        conv_block_conf=deepcopy(fixup),
        shortcut_block_conf=enc_shortcuts,
        vq_conf=vq_conf,
        n_pre_enc_layers=list(n_pre_enc_layers),
    )
    decoder_conf = None
    if all(nd == 0 for nd in level_downs[1:]) and shortcut_mode == "same":
        # decoder shortcut i: instantiated with out_channels = widths[i+1] (== widths[i] here)
        dec_shortcuts = [{**deepcopy(fixup), "mode": "same", "in_channels": widths[i + 1]}
                         for i in range(n_levels - 1)]
        decoder_conf = dict(
            _target_="vq_ae.model.Decoder", _recursive_=False,
            n_enc_layers=n_levels,
            stem_conf=same2d(in_channels=stem_channels, out_channels=in_channels),
            up_block_conf=[up(nd) for nd in level_downs],
            conv_block_conf=deepcopy(fixup),
            shortcut_block_conf=dec_shortcuts,
            n_post_enc_layers=list(n_pre_enc_layers),
        )
    return dict(
        _target_="vq_ae.model.VQAE", _recursive_=False,
        optim_conf=dict(_target_="torch.optim.AdamW", lr=1e-4),
        loss_f_conf=dict(_target_="torch.nn.modules.loss.HuberLoss", reduction="mean", delta=1.0),
        encoder_conf=encoder_conf,
        decoder_conf=decoder_conf,
    )


def up2d(**over) -> Dict[str, Any]:
    """conv_layer/up2d.yaml over convtranspose2d.yaml: ConvTranspose2d kernel 2, stride 2."""
    base = dict(_target_="torch.nn.ConvTranspose2d", in_channels=None, out_channels=None, kernel_size=2,
                stride=2, padding=0, output_padding=0, groups=1, bias=True, dilation=1,
                padding_mode="zeros")
    base.update(over)
    return base


def mbconv(expand_ratio: float = 4, batchnorm: bool = True, se: bool = True) -> Dict[str, Any]:
    """conf/model/layers/conv_block/mbconv.yaml:1-80 resolved (SiLU, BatchNorm2d, SELayer with bottleneck
    divisor 4; depthwise branch_conv2 whose ``groups`` the block fills in)."""
    nb = dict(bias=False)
    circ = dict(bias=False, padding_mode="circular")
    return dict(
        _target_="vq_ae.layers.conv_block.MBConv", _recursive_=False,
        in_channels=None, out_channels=None, mode=None, expand_ratio=expand_ratio,
        activation_conf=dict(_target_="torch.nn.SiLU"),
        batchnorm_conf=(dict(_target_="torch.nn.BatchNorm2d", num_features=None, eps=1e-05, momentum=0.1,
                             affine=True, track_running_stats=True) if batchnorm else None),
        se_conf=(dict(_target_="vq_ae.layers.misc.SELayer", in_channels=None, out_channels=None,
                      bottleneck_divisor=4) if se else None),
        conv_conf=dict(
            down=dict(branch_conv1=proj2d(**nb), branch_conv2={**down2d(**circ), "groups": None},
                      branch_conv3=proj2d(**nb), skip_conv=down2d(**circ)),
            up=dict(branch_conv1=proj2d(**nb), branch_conv2={**up2d(**nb), "groups": None},
                    branch_conv3=proj2d(**nb), skip_conv=up2d(**nb)),
            same=dict(branch_conv1=proj2d(**nb), branch_conv2={**same2d(**circ), "groups": None},
                      branch_conv3=proj2d(**nb), skip_conv=proj2d(**nb)),
            out=dict(branch_conv1=proj2d(**nb), branch_conv2={**out2d(**circ), "groups": None},
                     branch_conv3=proj2d(**nb), skip_conv=out2d(**nb)),
        ),
    )


def compose_efficientnetv2_conf(n_down: int = 3, n_enc_layers_trunk: int = 3, **kwargs) -> Dict[str, Any]:
    """conf/model/{encoder,decoder}/efficientnetv2.yaml: the shipped tree with every conv block swapped
    for MBConv (pyramid blocks and trunks; stems and quantiser unchanged)."""
    conf = compose_vqae_conf(n_down=n_down, n_enc_layers_trunk=n_enc_layers_trunk, **kwargs)
    for part, pyramid in (("encoder_conf", "down_block_conf"), ("decoder_conf", "up_block_conf")):
        conf[part][pyramid]["conv_conf"] = mbconv()
        conf[part]["conv_block_conf"] = mbconv()
    return conf
