"""Model of the peer-memory halo exchange as the SpMV kernels run it (csrc/spmv_device.cuh), on the
CPU: P ranks, each a small state machine (wait for the acknowledgement of SpMV hseq-2, push its
entries into the peers' landing buffer hseq & 1, consume the peers' entries, acknowledge), stepped
in random order so that ranks drift apart as far as the protocol lets them.  Two packagings of the
landing buffers are modelled:

  flags : bare values, published by ONE flag per (buffer, source) after all of them are written --
          the measured round-1 protocol;
  ll    : every entry carries the sequence number of its SpMV (payload+flag records), nothing is
          published, the consumer polls each entry -- the opt-in SIGB_HALO_LL=1 variant.

Checked for both: no schedule deadlocks, and every value a consumer reads is the one its producer
computed for THAT SpMV (never a stale or a too-new one), also when entries of one push land in any
order."""
import numpy as np
import pytest


def run(P, nsteps, nent, ll, rng):
    value = lambda src, h, k: src * 1_000_000 + h * 1_000 + k             # what rank src sends as entry k of SpMV h
    # landing[dst][buf][src] = list of (value, seq) per entry ; hflag[dst][buf][src] ; ack[src][dst]
    landing = [[[[(None, 0)] * nent for _ in range(P)] for _ in range(2)] for _ in range(P)]
    hflag = [[[0] * P for _ in range(2)] for _ in range(P)]
    ack = [[0] * P for _ in range(P)]                                     # ack[src][dst]: last SpMV dst finished reading
    st = [{"h": 1, "phase": "ack", "pending": [], "todo": None, "unpublished": False} for _ in range(P)]
    done = 0
    idle = 0
    while done < P:
        progressed = False
        for r in rng.permutation(P):
            s = st[r]
            if s["h"] > nsteps:
                continue
            h, buf = s["h"], s["h"] & 1
            peers = [q for q in range(P) if q != r]
            if s["phase"] == "ack":                       # producer: buffer `buf` of every peer must have been read
                if h <= 2 or all(ack[r][q] >= h - 2 for q in peers):
                    s["pending"] = [(q, k) for q in peers for k in range(nent)]
                    rng.shuffle(s["pending"])             # stores land in any order
                    s["phase"] = "push"
                    progressed = True
            elif s["phase"] == "push":
                for _ in range(int(rng.integers(1, 4))):  # a few stores per step
                    if s["pending"]:
                        q, k = s["pending"].pop()
                        landing[q][buf][r][k] = (value(r, h, k), h if ll else 0)
                        progressed = True
                if not s["pending"]:
                    if not ll:
                        for q in peers:                   # fence, then ONE flag per destination
                            hflag[q][buf][r] = h
                    s["todo"] = [(q, k) for q in peers for k in range(nent)]
                    s["phase"] = "consume"
                    progressed = True
            elif s["phase"] == "consume":
                rest = []
                for q, k in s["todo"]:
                    v, seq = landing[r][buf][q][k]
                    ready = (seq == h) if ll else (hflag[r][buf][q] >= h)
                    if ready:
                        assert v == value(q, h, k), (r, q, h, k, v)      # never stale, never too new
                        progressed = True
                    else:
                        rest.append((q, k))
                s["todo"] = rest
                if not rest:
                    for q in peers:
                        ack[q][r] = h                     # tell every source its buffer was read
                    s["h"] += 1
                    s["phase"] = "ack"
                    if s["h"] > nsteps:
                        done += 1
        idle = 0 if progressed else idle + 1
        assert idle < 3, "deadlock: no rank can make progress"


@pytest.mark.parametrize("ll", [False, True], ids=["flags", "ll"])
@pytest.mark.parametrize("P", [2, 3, 8])
def test_halo_protocol_model(P, ll):
    rng = np.random.default_rng(100 * P + ll)
    for _ in range(5):
        run(P, nsteps=9, nent=int(rng.integers(1, 5)), ll=ll, rng=rng)
