"""``liteasr.nets``-compatible modules (same constructor signatures, same parameter names / state_dict schema) whose
forward/backward run on liblasr's sm_100a kernels.  Only the pieces on the U2 + hybrid-CTC hot path exist
(SURVEY.md section 8a): Conformer encoder with relative-position attention, Transformer decoder, CTC head."""
from .ctc import CTC  # noqa: E402,F401
from .transformer_decoder import TransformerDecoder  # noqa: E402,F401
from .transformer_encoder import TransformerEncoder  # noqa: E402,F401
