"""Host-side model of the block-skewed hand-off of `resblock_skew_kernel`: two sequential agents - the MMA
issuer and the epilogue warps - with exactly the waits and arrivals of the kernel.  The model checks what the kernel relies
on: the protocol terminates for every tile shape (no wait depends on a later arrival of the same round), the two hazards it
guards (an epilogue never rewrites operand rows an MMA still reads, an MMA never reads rows before their epilogue) hold,
and the schedule beats the per-conv hand-off of `resblock_tc_kernel` whenever there are at least three row blocks."""
import itertools

import pytest

N_CONV = 6


def simulate(MB, m_clk, e_clk, load_clk):
    """Returns (makespan, mma[(b,q)] = (start, end), epi[(b,q)] = (start, end)).  a_ready[b] / acc_ready[b] are completion
    times per round, like the phases of the kernel's mbarriers."""
    a_ready = {}          # (b, round) -> time; round 0 = loader, round q+1 = epilogue of conv q
    acc_ready = {}        # (b, q) -> time
    mma, epi = {}, {}
    # workers, round 0: loader handles two row blocks per batch and arrives for both afterwards
    t_w = 0.0
    for b0 in range(0, MB, 2):
        t_w += 2 * load_clk
        a_ready[(b0, 0)] = a_ready[(b0 + 1, 0)] = t_w
    # the two agents run concurrently; emulate with per-agent clocks and a fixed-point iteration over the dependency order
    t_i = 0.0             # issuer clock (issue is asynchronous: the tensor pipe executes in issue order)
    pipe_free = 0.0
    pending_w = [(q, b) for q in range(N_CONV) for b in range(MB)]
    pending_i = [(q, b) for q in range(N_CONV) for b in range(MB)]
    progress = True
    while (pending_w or pending_i) and progress:
        progress = False
        if pending_i:
            q, b = pending_i[0]
            deps = [(nb, q) for nb in (b - 1, b, b + 1) if 0 <= nb < MB]
            if all(d in a_ready for d in deps):
                start = max(t_i, pipe_free, max(a_ready[d] for d in deps))
                end = start + m_clk
                mma[(b, q)] = (start, end)
                acc_ready[(b, q)] = end
                pipe_free, t_i = end, start        # the issuer only blocks on its waits, not on the MMA itself
                pending_i.pop(0)
                progress = True
        if pending_w:
            q, b = pending_w[0]
            deps = [(b, q)] + ([(b + 1, q)] if (q < N_CONV - 1 and b + 1 < MB) else [])
            if all(d in acc_ready for d in deps):
                start = max(t_w, max(acc_ready[d] for d in deps))
                end = start + e_clk
                epi[(b, q)] = (start, end)
                if q < N_CONV - 1:
                    a_ready[(b, q + 1)] = end
                t_w = end
                pending_w.pop(0)
                progress = True
    assert not pending_w and not pending_i, "protocol deadlocked"
    return max(t_w, pipe_free), mma, epi


@pytest.mark.parametrize("MB", [2, 4, 8])
@pytest.mark.parametrize("m_clk,e_clk", [(240, 500), (880, 500), (2100, 650), (100, 100)])
def test_protocol_terminates_and_respects_hazards(MB, m_clk, e_clk):
    makespan, mma, epi = simulate(MB, m_clk, e_clk, load_clk=700)
    for q, b in itertools.product(range(N_CONV), range(MB)):
        if q < N_CONV - 1:
            # WAR: the epilogue of (b, q) rewrites operand rows that MMA(b+1, q) (and, in issue order, MMA(b-1, q)) still read
            for nb in (b - 1, b + 1):
                if 0 <= nb < MB:
                    assert epi[(b, q)][0] >= mma[(nb, q)][1] - 1e-9
            # RAW: MMA(b, q+1) reads rows of blocks b-1, b, b+1 written by the epilogues of conv q
            for nb in (b - 1, b, b + 1):
                if 0 <= nb < MB:
                    assert mma[(b, q + 1)][0] >= epi[(nb, q)][1] - 1e-9
        # accumulator reuse: MMA(b, q+1) overwrites acc[b] only after the epilogue of (b, q) has drained it
        if q < N_CONV - 1:
            assert mma[(b, q + 1)][0] >= epi[(b, q)][1] - 1e-9
    assert makespan > 0


@pytest.mark.parametrize("MB,m_clk,e_clk", [(4, 240, 500), (4, 880, 500), (8, 140, 260), (4, 2100, 650)])
def test_skewed_schedule_beats_per_conv_handoff(MB, m_clk, e_clk):
    load = 700
    skew, _, _ = simulate(MB, m_clk, e_clk, load)
    serial = MB * load + N_CONV * MB * (m_clk + e_clk)       # resblock_tc_kernel: MMA of all blocks, then epilogue of all blocks
    # lower bound: the busier of the two agents (the first MMA needs two loaded row blocks; the loader overlaps the rest)
    bound = max(2 * load + N_CONV * MB * m_clk, MB * load + N_CONV * MB * e_clk)
    assert skew < 0.9 * serial
    assert skew >= bound - 1e-9
    assert skew <= bound + N_CONV * 2 * (m_clk + e_clk) + 1  # at most about two blocks of pipeline fill per conv
