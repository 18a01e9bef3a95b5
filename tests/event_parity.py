"""Instrument for the north-star requirement "the thresholded MIDI event list is bit-exact for frames whose probabilities
are not within tolerance of a threshold" (BASELINE.json; rust-plugins/src/common.rs:47-144).

extract_events is, per key, a state machine whose every branch is one of five comparisons:

    idle:     p[f] > 0.5                                   (activation,      common.rs:137)
    playing:  p[f] < 0.1                                   (release,         common.rs:81)
              p[f] < p[f+1]                                (decide on the local maximum, common.rs:114)
              p[f] > 0.4                                   (re-attack level, common.rs:119)
              mean(p[f..f+6]) - mean(p[f-6..f]) > 0.1      (re-attack rise,  common.rs:93-112; needed only if p[f] > 0.4)

`decision_margins` replays the machine on the REFERENCE probabilities and records, per key, the smallest distance of any
comparison it actually evaluated from flipping: m1 for comparisons of one probability with a constant, m2 for comparisons
of two probabilities / two means.  If another probability track differs from the reference by at most d everywhere on that
key, a one-probability comparison can only flip when m1 <= d and a two-sided one only when m2 <= 2 d (+ f32 rounding of the
six-term sums); by induction over frames, a key with m1 > d and m2 > 2 d + 1e-5 goes through exactly the same states, so its
event list MUST be identical.  `check_event_parity` asserts exactly that, with d measured per key and bounded by the stated
tolerance -- nothing is excused by a global clause, and the number of keys for which the assertion had force is returned so
that callers can pin it.
"""
import numpy as np

f32 = np.float32


def decision_margins(probs):
    """probs [F, K] float32 -> (events, m1[K], m2[K]); events as oracle.events.extract_events returns them."""
    probs = np.asarray(probs, dtype=np.float32)
    F, K = probs.shape
    m1 = np.full(K, np.inf)
    m2 = np.full(K, np.inf)
    events = []
    for key in range(K):
        p = probs[:, key]
        started = None
        for f in range(F):
            cur = p[f]
            if started is None:
                m1[key] = min(m1[key], abs(float(cur) - 0.5))
                if cur > f32(0.5):
                    started = f
                continue
            m1[key] = min(m1[key], abs(float(cur) - 0.1))
            if cur < f32(0.1):
                events.append((started, key, max(f - started, 1), 7))
                started = None
                continue
            rise_margin = None
            should = False
            if f32(f) - f32(started) > 5.0:
                prev = f32(0.0)
                for i in range(f - 6, f):
                    prev = f32(prev + p[i])
                prev = f32(prev / f32(6.0))
                nxt = f32(0.0)
                for i in range(f, min(f + 6, F)):
                    nxt = f32(nxt + p[i])
                nxt = f32(nxt / f32(6.0))
                should = f32(nxt - prev) > f32(0.1)
                rise_margin = abs(float(f32(nxt - prev)) - 0.1)
            if f < F - 1:
                m2[key] = min(m2[key], abs(float(cur) - float(p[f + 1])))
                if cur < p[f + 1]:
                    continue
            m1[key] = min(m1[key], abs(float(cur) - 0.4))
            if cur > f32(0.4):
                if rise_margin is not None:
                    m2[key] = min(m2[key], rise_margin)
                if should:
                    events.append((started, key, max(f - 1 - started, 1), 7))
                    started = f
        if started is not None:
            events.append((started, key, max(F - started, 1), 7))
    events.sort()
    return events, m1, m2


def decided_keys(p_ref, p_other):
    """Boolean [K]: keys whose event list is forced to be identical between the two probability tracks (see module doc)."""
    _, m1, m2 = decision_margins(p_ref)
    d = np.abs(np.asarray(p_other, np.float64) - np.asarray(p_ref, np.float64)).max(axis=0)
    return (m1 > d) & (m2 > 2.0 * d + 1e-5)


def check_event_parity(p_ref, p_other, events_other, tol):
    """Asserts |p_other - p_ref| <= tol everywhere and, for every decided key, that `events_other` (the product's event list
    for p_other) restricted to that key equals the reference machine's list on p_ref.  Returns (n_decided, n_identical)."""
    p_ref = np.asarray(p_ref, np.float32)
    p_other = np.asarray(p_other, np.float32)
    worst = float(np.abs(p_other.astype(np.float64) - p_ref.astype(np.float64)).max())
    assert worst <= tol, f"probabilities differ by {worst:.3e} > tolerance {tol:.1e}"
    ev_ref, m1, m2 = decision_margins(p_ref)
    d = np.abs(p_other.astype(np.float64) - p_ref.astype(np.float64)).max(axis=0)
    decided = (m1 > d) & (m2 > 2.0 * d + 1e-5)
    by_key_o, by_key_r = {}, {}
    for e in events_other:
        by_key_o.setdefault(int(e[1]), []).append(tuple(int(x) for x in e))
    for e in ev_ref:
        by_key_r.setdefault(int(e[1]), []).append(tuple(int(x) for x in e))
    same = 0
    for key in range(p_ref.shape[1]):
        a, b = by_key_o.get(key, []), by_key_r.get(key, [])
        if a == b:
            same += 1
        elif decided[key]:
            raise AssertionError(f"key {key}: event lists differ although every decision on this key has a margin above the "
                                 f"measured probability difference {d[key]:.2e} (m1 {m1[key]:.2e}, m2 {m2[key]:.2e}): {a} vs {b}")
    return int(decided.sum()), same
