"""Model check of the peer-to-peer top-k exchange of comm.cu (sdb_knn_sharded_*), on the CPU.

The exchange cannot be exercised without several GPUs, and its failure mode is a hang, so the PROTOCOL (not the CUDA
code) is restated here as a small discrete-event model and run under thousands of random interleavings:

  per rank:  two batch streams (ticket slot parity), 4 ticket slots, batches submitted with two in flight;
  per batch: push  = S "CTAs" per peer: wait for the peer's acknowledgement of the slot's previous use, copy a share of
                     the block into the peer's gather slot, count the arrival; the LAST arrival publishes the batch's
                     sequence number in the peer's flag word                         (exch_push_kernel)
             wait  = until every rank's flag of this slot has reached the sequence number  (exch_wait_kernel)
             merge = reads the slot's blocks (here: checks that they all belong to this batch)
             ack   = tells every peer the slot may be overwritten                     (exch_ack_kernel)

Properties checked: every schedule terminates (no deadlock) and every merge reads exactly its own batch's blocks.
The model also reproduces the bug that hung the 50-step 8-GPU run of round 2 -- ONE arrival counter per peer, shared
by the pushes of both streams -- which is what the per-(slot, peer) counters fixed.
"""
import random

N_TICKETS = 4


class Rank:
    def __init__(self, r, n_ranks, n_cta):
        self.r = r
        self.gather = [[[None] * n_cta for _ in range(n_ranks)] for _ in range(N_TICKETS)]  # [slot][src] -> share per CTA
        self.flags = [[0] * n_ranks for _ in range(N_TICKETS)]
        self.acks = [[0] * n_ranks for _ in range(N_TICKETS)]
        self.ctr = {}
        self.streams = [[], []]      # FIFO of ops per stream; an op is a list of independent sub-ops (CTAs)
        self.busy = [False] * N_TICKETS
        self.done = set()            # batches whose merge + ack have run
        self.seq = 0
        self.slot_seq = [0] * N_TICKETS
        self.next_batch = 0          # host program state
        self.pending = []            # (batch, slot) submitted and not yet waited for
        self.repairing = set()


def simulate(n_ranks, n_batches, depth, n_cta, per_slot_counters, rng, max_steps=400_000, repair_every=0):
    """Sub-ops return True (finished), None (advanced, not finished) or False (blocked: no state change)."""
    ranks = [Rank(r, n_ranks, n_cta) for r in range(n_ranks)]
    errors = []

    def submit(rk, b, slot=None, tag=None):
        if slot is None:
            slot = rk.busy.index(False)
            rk.busy[slot] = True
            rk.pending.append((b, slot))
        tag = b if tag is None else tag  # what the blocks of this exchange carry / what marks it complete
        rk.seq += 1
        seq, need = rk.seq, rk.slot_seq[slot]
        rk.slot_seq[slot] = seq
        st = rk.streams[slot & 1]

        def cta(p, s):
            state = {"pc": 0}

            def step():
                peer = ranks[p]
                if state["pc"] == 0:  # wait for the peer's acknowledgement of the slot's previous use
                    if rk.acks[slot][p] < need:
                        return False
                    state["pc"] = 1
                    return None
                if state["pc"] == 1:  # copy this CTA's share
                    peer.gather[slot][rk.r][s] = tag
                    state["pc"] = 2
                    return None
                key = (slot, p) if per_slot_counters else p  # arrival counter
                old = rk.ctr.get(key, 0)
                rk.ctr[key] = old + 1
                if old == n_cta - 1:
                    rk.ctr[key] = 0
                    peer.flags[slot][rk.r] = seq
                return True
            return step

        def wait_op():
            return True if all(f >= seq for f in rk.flags[slot]) else False

        def merge_op():
            for src in range(n_ranks):
                if rk.gather[slot][src] != [tag] * n_cta:
                    errors.append((rk.r, tag, slot, src, list(rk.gather[slot][src])))
            return True

        def ack_op():
            for p in range(n_ranks):
                ranks[p].acks[slot][rk.r] = seq
            rk.done.add(tag)
            return True

        st.append([cta(p, s) for p in range(n_ranks) for s in range(n_cta)])  # one kernel, independent CTAs
        st.append([wait_op])
        st.append([merge_op])
        st.append([ack_op])

    def host_step(rk):
        """bench.py's loop: submit; once `depth` batches are in flight, wait for the oldest.  False = blocked."""
        if rk.next_batch < n_batches and len(rk.pending) < depth:
            submit(rk, rk.next_batch)
            rk.next_batch += 1
            return True
        if rk.pending and rk.pending[0][0] in rk.done:
            b, slot = rk.pending[0]
            if repair_every and b % repair_every == 0 and (b, "repair") not in rk.done:
                # every rank saw a non-zero header: a second exchange through the SAME slot (finish_all's repair round),
                # enqueued by the host inside the wait call
                if (b, "repair") not in rk.repairing:
                    rk.repairing.add((b, "repair"))
                    submit(rk, b, slot=slot, tag=(b, "repair"))
                    return True
                return False
            rk.pending.pop(0)
            rk.busy[slot] = False
            return True
        return False

    for _ in range(max_steps):
        actions = [("host", rk.r) for rk in ranks]
        for rk in ranks:
            for si, st in enumerate(rk.streams):
                if st:
                    actions += [("op", rk.r, si, j) for j in range(len(st[0]))]
        rng.shuffle(actions)
        progressed = False
        for act in actions:  # take the first action (in random order) that is not blocked
            if act[0] == "host":
                if host_step(ranks[act[1]]):
                    progressed = True
                    break
                continue
            _, r, si, j = act
            st = ranks[r].streams[si]
            res = st[0][j]()
            if res is False:
                continue
            if res is True:
                st[0].pop(j)
                if not st[0]:
                    st.pop(0)
            progressed = True
            break
        if not progressed:
            finished = all(rk.next_batch == n_batches and not rk.pending and not any(rk.streams) for rk in ranks)
            return ("done" if finished else "deadlock"), errors
    return "timeout", errors


def run_many(per_slot, n_runs, seed, **kw):
    out = {"done": 0, "deadlock": 0, "timeout": 0, "corrupt": 0}
    for i in range(n_runs):
        rng = random.Random(seed * 100003 + i)
        status, errors = simulate(per_slot_counters=per_slot, rng=rng, **kw)
        out[status] += 1
        out["corrupt"] += 1 if errors else 0
    return out


def test_exchange_protocol_never_deadlocks_and_never_mixes_batches():
    for n_ranks, n_cta, depth, repair in ((2, 2, 2, 0), (3, 2, 2, 0), (4, 1, 2, 0), (3, 3, 3, 0), (2, 2, 4, 0), (3, 2, 2, 3), (2, 2, 2, 1)):
        res = run_many(True, 150, seed=n_ranks * 10 + n_cta, n_ranks=n_ranks, n_batches=9, depth=depth, n_cta=n_cta,
                       repair_every=repair)
        assert res["done"] == 150 and res["corrupt"] == 0, (n_ranks, n_cta, depth, repair, res)


def test_model_reproduces_the_shared_counter_hang():
    # one arrival counter per peer, shared by the two streams' pushes: some interleavings publish a flag early (the merge
    # reads a half-copied block) and leave the other push's flag unpublished for ever
    res = run_many(False, 400, seed=7, n_ranks=2, n_batches=9, depth=2, n_cta=2)
    assert res["deadlock"] + res["corrupt"] > 0, res
