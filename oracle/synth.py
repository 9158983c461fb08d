"""Synthetic inputs now live in the product package (``msa_b200.synth``: bench.py's native arm
needs them and may not import ``oracle/``); the oracle and the tests keep this name."""
import msa_b200  # noqa: F401  (registers the package alias)
from msa_b200.synth import *  # noqa: F401,F403
from msa_b200.synth import (AUDIO_DIM, FACE_DIM, HIDDEN_DIM, OUT_DIM, SAMPLE_RATE, SEGMENT_SAMPLES,  # noqa: F401
                            SEGMENT_SECONDS, TEXT_DIM)
