"""Drop-in counterparts of the reference's ``vq_ae.layers`` package."""
from . import conv, conv_block, vq  # noqa: F401
