"""Generate tests/golden/e2e_golden.npz: the UNMODIFIED reference, end to end, on 10,240 synthetic segments.

TEST INFRASTRUCTURE.  Runs only in the build container (needs /root/reference).  For every segment the chain of
/root/reference/src/processors/streaming_processor.py:250-320 is executed with the reference's own code:

    AudioAnalyzer._analyze_* / _calculate_*   (src/analyzers/audio_analyzer.py:152-329, imported by file path)
    -> AudioFeatureNormalizer.normalize       (src/utils/normalization.py:26-44)
    -> audio row = LN slices ++ quality4, torch.nan_to_num(nan=0)           (streaming_processor.py:250-268, 295-298)
    -> AdvancedFusionModel.forward(face, audio, text | None)["fused"]       (src/models/fusion_model.py:131-190)

so that tests/test_gpu_e2e.py can compare the GPU pipeline's rows, logits and arg-max with the reference's, NOT with
an oracle that was fed the GPU's own feature rows.  Four blocks of 2,560 segments cover the modes the north star names:

    block 0  reference as is (mono intensity = NaN -> all-NaN LayerNorm row), uniform 1/8 emotion, 3-modal fusion
    block 1  reference as is, injected non-uniform emotion embedding, face + audio fusion
    block 2  finite row (_analyze_intensity takes its own except-branch default, zeros), injected emotion, 3-modal
    block 3  finite row, uniform emotion, face + audio

Stated deviation (as everywhere in this repo): the fusion model runs in eval mode (the reference forgets .eval()).
Inputs are regenerated from seeds by multimodal-sentiment-analyzer_b200/synth.py (numpy PCG64: same bytes everywhere).

    python -m oracle.make_golden_e2e            # ~1 minute on 8 cores
"""
from __future__ import annotations

import os
import sys
import warnings
from multiprocessing import Pool

import numpy as np

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "e2e_golden.npz")
BLOCK = 2560
CHUNK = 256                      # segments per synthesis call / worker task
WAVE_SEED, EMO_SEED, FACE_SEED, TEXT_SEED, WEIGHT_SEED = 50_000, 60_000, 70_000, 80_000, 4321
BLOCKS = [  # (strict NaN, uniform emotion, three modalities)
    (True, True, True),
    (True, False, False),
    (False, False, True),
    (False, True, False),
]


def chunk_inputs(block: int, chunk: int):
    """Inputs of chunk `chunk` of block `block` (shared by this generator and the GPU test)."""
    from oracle import synth
    strict, uniform, three = BLOCKS[block]
    k = block * (BLOCK // CHUNK) + chunk
    pcm = synth.fast_segments_pcm(WAVE_SEED + k, CHUNK)
    emo = None if uniform else synth.emotion_probs(EMO_SEED + k, CHUNK)
    face = synth.face_rows(FACE_SEED + k, CHUNK)
    text = synth.text_rows(TEXT_SEED + k, CHUNK) if three else None
    return pcm, emo, face, text


_state = {}


def _worker_init():
    import torch
    torch.set_num_threads(1)
    from oracle.make_golden import load_reference
    from oracle import synth
    audio_mod, fusion_mod = load_reference()
    ana = audio_mod.AudioAnalyzer(device="cpu")
    model = fusion_mod.AdvancedFusionModel(device="cpu")
    sd = synth.fusion_state(WEIGHT_SEED, trained_like=True)
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    model.eval()
    _state.update(ana=ana, model=model, torch=torch)


class _InjectedSER:
    """Stands in for the out-of-scope wav2vec2 classifier: returns the injected [1, 8] embedding the way
    classify_batch returns out_prob (audio_analyzer.py:155-156)."""

    def __init__(self):
        self.row = None

    def classify_batch(self, w):
        if self.row is None:
            raise RuntimeError("no embedding: _analyze_emotion takes its uniform fallback")
        return self.row, None, None, None


def _run_chunk(args):
    block, chunk = args
    torch, ana, model = _state["torch"], _state["ana"], _state["model"]
    from oracle import synth
    strict, uniform, three = BLOCKS[block]
    pcm, emo, face, text = chunk_inputs(block, chunk)
    ser = _InjectedSER()
    ana.emotion_model = ser
    if strict:
        ana.__dict__.pop("_analyze_intensity", None)
    else:
        # the reference's own failure default of _analyze_intensity (audio_analyzer.py:199-201)
        ana._analyze_intensity = lambda w: torch.zeros(1, 1)
    rows = np.zeros((CHUNK, 31), np.float32)
    logits = np.zeros((CHUNK, 7), np.float32)
    with warnings.catch_warnings(), torch.no_grad():
        warnings.simplefilter("ignore")
        for i in range(CHUNK):
            w = torch.from_numpy(synth.pcm_to_f32(pcm[i]))[None, :]
            ser.row = None if emo is None else torch.from_numpy(emo[i])[None, :]
            # body of AudioAnalyzer.analyze (audio_analyzer.py:83-147) without the file read
            feats = torch.cat([ana._analyze_emotion(w), ana._analyze_pitch(w), ana._analyze_intensity(w), ana._analyze_timbre(w),
                               ana._analyze_speech_rate(w), ana._analyze_rhythm(w)], dim=1)
            feats = ana.normalizer.normalize(feats)
            q = torch.tensor([ana._calculate_audio_quality(w), ana._calculate_signal_noise_ratio(w), ana._calculate_clarity(w),
                              ana._calculate_consistency(w)]).float()[None, :]
            # streaming_processor.py:250-268, 295-298
            audio_row = torch.nan_to_num(torch.cat([feats[:, :27].float(), q], dim=1), nan=0.0)
            out = model(torch.from_numpy(face[i])[None, :], audio_row, None if text is None else torch.from_numpy(text[i])[None, :])
            rows[i] = audio_row.numpy()[0]
            logits[i] = out["fused"].numpy()[0]
    return block, chunk, rows, logits


def main():
    sys.dont_write_bytecode = True
    tasks = [(b, c) for b in range(len(BLOCKS)) for c in range(BLOCK // CHUNK)]
    n = len(BLOCKS) * BLOCK
    rows = np.zeros((n, 31), np.float32)
    logits = np.zeros((n, 7), np.float32)
    with Pool(processes=os.cpu_count(), initializer=_worker_init) as pool:
        for block, chunk, r, l in pool.imap_unordered(_run_chunk, tasks):
            o = block * BLOCK + chunk * CHUNK
            rows[o:o + CHUNK], logits[o:o + CHUNK] = r, l
            print(f"block {block} chunk {chunk} done", flush=True)
    top2 = np.sort(logits, axis=1)[:, -2:]
    np.savez_compressed(OUT, audio_rows=rows, logits=logits, argmax=logits.argmax(1).astype(np.int8),
                        top2_gap=(top2[:, 1] - top2[:, 0]).astype(np.float32),
                        meta=np.array([BLOCK, CHUNK, WAVE_SEED, EMO_SEED, FACE_SEED, TEXT_SEED, WEIGHT_SEED]))
    print(OUT, os.path.getsize(OUT), "bytes; min top-2 logit gap", float((top2[:, 1] - top2[:, 0]).min()))


if __name__ == "__main__":
    main()
