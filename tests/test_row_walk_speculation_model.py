"""Why the speculative row walk (barycentric.cu, row_walk_kernel) is exact, as a model on the CPU.  The reference walks a
grid row column by column, every target from the triangle of the previous one (interpolation.f90:366-373): the state
after a column is a function of the state before it and of the column, nothing else.  The kernel cuts the chain
into 32 pieces, walks piece L from a GUESSED state while recording the state after every column, then repairs the
pieces in order: the true chain is continued into a piece only until its state equals the recorded one -- from there
on the recorded chain IS the true chain.  Any transition function, any guesses."""
import numpy as np


def speculative_chain(step, start, guesses, ncols, pieces=32):
    piece = (ncols + pieces - 1) // pieces
    rec = np.empty(ncols, dtype=np.int64)
    for lane in range(pieces):                                   # phase A: all pieces "in parallel"
        lo, hi = lane * piece, min((lane + 1) * piece, ncols)
        if lo >= ncols:
            break
        st = start if lane == 0 else guesses[lane]
        for c in range(lo, hi):
            st = step(st, c)
            rec[c] = st
    repaired = 0
    for lo in range(piece, ncols, piece):                        # phase B: one lane, in order
        st = rec[lo - 1]                                         # the true state after column lo - 1
        for c in range(lo, min(lo + piece, ncols)):
            st = step(st, c)
            if st == rec[c]:
                break                                            # met the recorded chain
            rec[c] = st
            repaired += 1
    return rec, repaired


def test_speculative_pieces_plus_repair_equal_the_serial_chain():
    rng = np.random.default_rng(1)
    for trial in range(300):
        nstates = int(rng.integers(2, 40)); ncols = int(rng.integers(64, 700))
        # a random transition table; `sticky` makes chains from different starts merge sooner or later (or never)
        table = rng.integers(0, nstates, (ncols, nstates))
        sticky = rng.uniform() < 0.5
        if sticky:
            for c in range(ncols):
                table[c, rng.integers(0, nstates, nstates // 2 + 1)] = rng.integers(0, nstates)
        step = lambda st, c: int(table[c, st])
        start = int(rng.integers(0, nstates))
        true = np.empty(ncols, dtype=np.int64)
        st = start
        for c in range(ncols):
            st = step(st, c)
            true[c] = st
        rec, _ = speculative_chain(step, start, rng.integers(0, nstates, 32), ncols)
        assert np.array_equal(rec, true), trial


def test_good_guesses_need_no_repair():
    ncols = 640
    step = lambda st, c: (st * 7 + c) % 101
    true = []
    st = 5
    for c in range(ncols):
        st = step(st, c)
        true.append(st)
    piece = ncols // 32
    guesses = [0] + [true[l * piece - 1] for l in range(1, 32)]
    rec, repaired = speculative_chain(step, 5, guesses, ncols)
    assert np.array_equal(rec, true) and repaired == 0
