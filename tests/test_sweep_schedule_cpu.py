"""The host-side scheduler of the kinematics kernel's packed tangent sweep (hippopt_b200/csrc/sweep_schedule.h),
checked on the CPU through hb_debug_sweep_schedule: every (direction, relevant body) step exactly once, chains stay
on one lane in increasing rounds, a slot is read only after every chain that feeds it has flushed, no two flushes
into one slot in the same round, and -- typed rounds -- the rounds the kernel runs on its light path hold propagation
tasks only."""
import ctypes

import numpy as np
import pytest

KT = dict(L=0, P=5, D=10, VALID=1 << 15, START=1 << 16, IN_L=1 << 17, IN_P=1 << 18, LOAD=19, FLUSH=25, STORE=1 << 31)


def schedule(lib, parent, feet, chest, typed):
    nb = len(parent)
    tasks = np.zeros(32 * 32, dtype=np.int32)
    info = np.zeros(40, dtype=np.int32)
    par = np.asarray(parent, dtype=np.int32)
    i32p = ctypes.POINTER(ctypes.c_int32)
    rc = lib.hb_debug_sweep_schedule(nb, par.ctypes.data_as(i32p), feet[0], feet[1], chest, int(typed),
                                     tasks.ctypes.data_as(i32p), info.ctypes.data_as(i32p))
    assert rc == 0
    return tasks.view(np.uint32).reshape(32, 32), info


def expected_tasks(parent, feet):
    nb = len(parent)
    anc = lambda b: ([] if parent[b] < 0 else [parent[b]] + anc(parent[b]))  # noqa: E731
    sub = [{l for l in range(nb) if b == l or b in anc(l)} for b in range(nb)]
    out = {}
    for d in range(4 + nb - 1):
        ld = 0 if d < 4 else d - 3
        rel = set(sub[ld]) | set(anc(ld))
        for f in range(2):
            if feet[f] in sub[ld]:
                rel |= {feet[1 - f]} | set(anc(feet[1 - f]))
        for l in rel:
            out[(d, l)] = l in sub[ld]
    return out


@pytest.mark.parametrize("typed", [1, 2, 0])  # 1: shipped policy, 2: strictly homogeneous rounds, 0: untyped
def test_schedule_is_a_valid_execution_of_the_sweep(model, built_library, typed):
    parent = [int(v) for v in model.parent]
    feet = (model.frames["l_sole"][0], model.frames["r_sole"][0])
    chest = model.frames["chest"][0]
    tasks, info = schedule(built_library, parent, feet, chest, typed)
    n_rounds, n_slots, n_tasks, n_heavy, heavy_mask = (int(info[0]), int(info[1]), int(info[2]), int(info[3]),
                                                       int(np.uint32(info[4])))
    want = expected_tasks(parent, feet)
    assert n_tasks == len(want) == 352 and n_heavy == sum(want.values()) == 188
    assert 0 < n_rounds <= 32 and n_slots * 12 <= 474
    seen, flushed_at, loads, root_store = {}, {}, [], {}
    running = {}  # lane -> (direction, last body, last round)
    for r in range(n_rounds):
        kinds, flush_slots = set(), []
        for lane in range(32):
            t = int(tasks[r, lane])
            if not t & KT["VALID"]:
                continue
            l, p, d = (t >> KT["L"]) & 31, (t >> KT["P"]) & 31, (t >> KT["D"]) & 31
            assert (d, l) in want and (d, l) not in seen
            seen[(d, l)] = (r, lane)
            assert bool(t & KT["IN_L"]) == want[(d, l)]
            if l != 0:
                assert p == parent[l] and bool(t & KT["IN_P"]) == want.get((d, p), False)
            kinds.add(bool(t & KT["IN_L"]))
            if t & KT["START"]:
                assert lane not in running
            else:  # continues the chain that sits on this lane: same direction, the child it just left
                pd, pl, pr = running[lane]
                assert pd == d and parent[pl] == l and pr < r
            running[lane] = (d, l, r)
            ls, fs = (t >> KT["LOAD"]) & 63, (t >> KT["FLUSH"]) & 63
            if ls:
                loads.append((ls - 1, r, d))
            if fs:
                flush_slots.append(fs - 1)
                if t & KT["STORE"]:  # the root task reads its slot (the children's sums) and stores the totals back
                    root_store[fs - 1] = (r, d)
                else:
                    flushed_at.setdefault(fs - 1, []).append((r, d))
                del running[lane]  # the chain ends here
        assert len(flush_slots) == len(set(flush_slots)), f"two flushes into one slot in round {r}"
        if typed and kinds:
            # a LIGHT round (bit clear) holds propagation tasks only; a heavy round holds at least one sub-tree task
            assert bool((heavy_mask >> r) & 1) == (True in kinds), f"round {r}: heavy bit does not match its tasks"
            if typed == 2:
                assert len(kinds) == 1, f"round {r} mixes sub-tree and propagation tasks"
    assert not running and set(seen) == set(want)
    for slot, r, d in loads:  # a slot is consumed after every flush into it, and only by its own direction
        assert slot in flushed_at and all(fr < r and fd == d for fr, fd in flushed_at[slot])
    for d in range(27):
        rs = int(info[6 + d])
        assert root_store[rs][1] == d and all(fr < root_store[rs][0] for fr, _ in flushed_at.get(rs, []))
    if typed:
        light = [r for r in range(n_rounds) if not (heavy_mask >> r) & 1]
        print(f"typed schedule: {n_rounds} rounds, {n_rounds - len(light)} heavy + {len(light)} light")
        assert len(light) >= 4 and n_rounds - len(light) <= 10


def test_other_tree_shapes(built_library):
    """a chain robot and a star: the scheduler depends on the topology only"""
    chain = [-1] + list(range(0, 9))
    tasks, info = schedule(built_library, chain, (9, 5), 3, True)
    assert info[2] == len(expected_tasks(chain, (9, 5)))
    star = [-1, 0, 0, 0, 0, 1, 2, 3, 4]
    tasks, info = schedule(built_library, star, (5, 6), 7, True)
    assert info[2] == len(expected_tasks(star, (5, 6)))
