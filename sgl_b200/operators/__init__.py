"""Host-side mirror of the reference's ``sgl.operators`` package (same class names, arguments and errors); the
arithmetic runs in libsglb200.so on the B200."""
from . import graph_op, message_op  # noqa: F401
from .base_op import GraphOp, MessageOp  # noqa: F401
