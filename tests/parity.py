"""Token-identity checker shared by the GPU parity tests and ``__graft_entry__.smoke()``.

north_star asks for "identical greedy token sequences on fixed test clips".  The engine stores activations in bf16 at a
handful of points, so a decision whose top-2 logits are closer than that rounding can legitimately flip; everything else
must be identical.  The checker therefore (1) rebuilds the engine's full DECISION sequence (the argmax of every joint
evaluation, blanks included) from its tokens + frames, (2) walks it through the oracle teacher-forced
(``oracle.nemo_restated.greedy_follow``): a difference never ends the comparison, the oracle is put on the engine's path
and the walk goes on to the last frame, (3) fails on ANY difference whose oracle logit gap is >= ``tol`` (default 1e-2,
SURVEY.md A.6) and on more than ``max_near_ties`` sub-tolerance differences per clip.
"""
from __future__ import annotations

from typing import List, Sequence

NEAR_TIE_TOL = 1e-2
MAX_NEAR_TIES = 3


def decisions_from(tokens: Sequence[int], frames: Sequence[int], T: int, max_symbols: int, blank: int) -> List[int]:
    """Per frame: the tokens emitted there, then a blank unless the frame was left because max_symbols was reached."""
    out, i = [], 0
    for t in range(T):
        n = 0
        while i < len(tokens) and frames[i] == t:
            out.append(int(tokens[i])); i += 1; n += 1
        if n < max_symbols:
            out.append(blank)
    assert i == len(tokens), f"{len(tokens) - i} tokens carry frames outside [0, {T}) or out of order"
    return out


def check_decisions(tokens, frames, enc_oracle, sd, cfg, tag: str, tol: float = NEAR_TIE_TOL,
                    max_near_ties: int = MAX_NEAR_TIES, emulate: bool = True) -> int:
    """Engine tokens/frames of ONE clip vs the oracle's predictor + joint on the oracle's own encoder output
    (``enc_oracle``: [T, d_model]).  Returns the number of near-tie differences (0 = identical sequence)."""
    from oracle import nemo_restated as O
    T = enc_oracle.shape[0]
    got = decisions_from(list(tokens), list(frames), T, cfg.max_symbols, cfg.blank)
    r = O.greedy_follow(enc_oracle, sd, cfg, got, emulate=emulate)
    assert r.complete, f"{tag}: the decision sequence does not cover the clip's {T} frames exactly ({len(got)} decisions, {r.n_decisions} consumed)"
    for (i, t, k, k_or, gap) in r.gaps:
        print(f"{tag}: decision {i} (frame {t}): engine {k}, oracle {k_or}, oracle logit gap {gap:.3e}")
    clear = [g for g in r.gaps if not g[4] < tol]
    assert not clear, f"{tag}: {len(clear)} decisions differ from the oracle at a logit gap >= {tol:g}: {clear[:4]}"
    assert len(r.gaps) <= max_near_ties, f"{tag}: {len(r.gaps)} near-tie differences (> {max_near_ties}) in {len(got)} decisions"
    return len(r.gaps)
