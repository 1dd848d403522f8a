"""Model of the peer-memory halo exchange as the SpMV kernels run it (csrc/spmv_device.cuh), on the
CPU: P ranks, each a small state machine (wait for the acknowledgement of SpMV hseq-2, push its
entries into the peers' landing buffer hseq & 1, consume the peers' entries, acknowledge), stepped
in random order so that ranks drift apart as far as the protocol lets them.  Two packagings of the
landing buffers are modelled:

  flags : bare values, published by ONE flag per (buffer, source) after all of them are written --
          the measured round-1 protocol;
  ll    : every entry carries the sequence number of its SpMV (payload+flag records), nothing is
          published, the consumer polls each entry -- measured in round 2 and dropped from the library
          (no gain once dedicated communication CTAs took the push); the model stays as a record.

A third model runs the protocol at CTA granularity for operators without locality (HaloSync::push_all):
EVERY CTA of a rank first pushes its share of the send list, the CTA that takes the last push ticket
publishes the flags, every CTA then waits for the peers' flags before its (boundary) tiles, the CTA that
takes the last done ticket acknowledges -- CTAs of one rank progress independently inside a launch, a
rank's next SpMV starts when all its CTAs have finished (kernel boundary).

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


def run_push_all(P, C, nsteps, nent, rng):
    """CTA-granular model of the push-by-every-CTA exchange (spmv_device.cuh: halo_push called by all CTAs,
    push_ticket / hflag / done_ticket / ack)."""
    value = lambda src, h, k: src * 1_000_000 + h * 1_000 + k
    landing = [[[[None] * nent for _ in range(P)] for _ in range(2)] for _ in range(P)]   # [dst][buf][src][k]
    hflag = [[[0] * P for _ in range(2)] for _ in range(P)]
    ack = [[0] * P for _ in range(P)]                       # ack[src][dst]
    h_of = [1] * P                                          # SpMV a rank is running
    push_ticket = [0] * P
    done_ticket = [0] * P
    # the send list of a rank (peer, entry) is dealt round-robin to its CTAs
    share = lambda r, c: [(q, k) for i, (q, k) in enumerate((q, k) for q in range(P) if q != r for k in range(nent)) if i % C == c]
    cta = [[{"phase": "ack", "pending": []} for _ in range(C)] for _ in range(P)]
    finished = 0
    idle = 0
    while finished < P:
        progressed = False
        order = [(r, c) for r in range(P) for c in range(C)]
        rng.shuffle(order)
        for r, c in order:
            h = h_of[r]
            if h > nsteps:
                continue
            s, buf = cta[r][c], h & 1
            peers = [q for q in range(P) if q != r]
            if s["phase"] == "ack":
                if h <= 2 or all(ack[r][q] >= h - 2 for q in peers):
                    s["pending"] = share(r, c)
                    rng.shuffle(s["pending"])
                    s["phase"] = "push"
                    progressed = True
            elif s["phase"] == "push":
                for _ in range(int(rng.integers(1, 4))):
                    if s["pending"]:
                        q, k = s["pending"].pop()
                        landing[q][buf][r][k] = value(r, h, k)
                        progressed = True
                if not s["pending"]:
                    push_ticket[r] += 1
                    if push_ticket[r] == C:               # the last pusher publishes
                        push_ticket[r] = 0
                        for q in peers:
                            hflag[q][buf][r] = h
                    s["phase"] = "consume"
                    progressed = True
            elif s["phase"] == "consume":
                if all(hflag[r][buf][q] >= h for q in peers):
                    for q in peers:
                        for k in range(nent):
                            assert landing[r][buf][q][k] == value(q, h, k), (r, c, q, h, k)
                    done_ticket[r] += 1
                    s["phase"] = "done"
                    progressed = True
                    if done_ticket[r] == C:               # the last CTA acknowledges; the launch ends
                        done_ticket[r] = 0
                        for q in peers:
                            ack[q][r] = h
                        h_of[r] += 1
                        for cc in range(C):
                            cta[r][cc]["phase"] = "ack"
                        if h_of[r] > nsteps:
                            finished += 1
        idle = 0 if progressed else idle + 1
        assert idle < 3, "deadlock: no CTA can make progress"


@pytest.mark.parametrize("P,C", [(2, 3), (3, 4), (8, 2)])
def test_push_by_every_cta_model(P, C):
    rng = np.random.default_rng(7 * P + C)
    for _ in range(4):
        run_push_all(P, C, nsteps=7, nent=int(rng.integers(1, 6)), rng=rng)
