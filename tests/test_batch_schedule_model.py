"""Host-side model of the batch kernel's work-item schedule (csrc/batch.cu, VisitSeq).

The TMA, MMA and epilogue warps of every unit (CTA or cluster) walk the same sequence of visits, generated
without divisions after start(): work item = (tile block of R consecutive database tiles, query-tile group),
numbered tile-block-major; unit u takes items u, u + step, ...  This test restates the generator in Python,
line by line, and checks the property the kernel relies on: over all units every (database tile, query-tile
group) pair is visited exactly once, in ascending tile order per (unit, group), for strides that are the unit
count (normal schedule) or a multiple of the group count (pinned schedule: a unit never changes group)."""
import itertools
import random


class VisitSeq:
    def __init__(self, q_tiles, n_tiles, tile_block, cl, first_item, item_step):
        self.n_groups = (q_tiles + cl - 1) // cl
        self.n_tiles = n_tiles
        self.R = tile_block
        n_tb = (n_tiles + self.R - 1) // self.R
        self.n_items = n_tb * self.n_groups
        self.step = item_step
        self.step_tb = item_step // self.n_groups
        self.step_g = item_step - self.step_tb * self.n_groups
        self.item = min(first_item, self.n_items)
        self.tb = self.item // self.n_groups
        self.g = self.item - self.tb * self.n_groups
        self.r = 0
        self.nr = min(self.R, self.n_tiles - self.tb * self.R)

    def done(self):
        return self.item >= self.n_items

    def t(self):
        return self.tb * self.R + self.r

    def next(self):
        self.r += 1
        if self.r < self.nr:
            return
        self.item += self.step
        self.tb += self.step_tb
        self.g += self.step_g
        if self.g >= self.n_groups:
            self.g -= self.n_groups
            self.tb += 1
        self.r = 0
        self.nr = min(self.R, self.n_tiles - self.tb * self.R)


def visits_of_unit(q_tiles, n_tiles, R, cl, unit, n_units, visit_stride):
    step = visit_stride if visit_stride > 0 else n_units
    first = unit if unit < step else 1 << 60
    seq = VisitSeq(q_tiles, n_tiles, R, cl, first, step)
    out = []
    while not seq.done():
        assert 0 <= seq.t() < n_tiles and 0 <= seq.g < seq.n_groups
        assert seq.item == seq.tb * seq.n_groups + seq.g          # the incremental walk never drifts
        out.append((seq.t(), seq.g))
        seq.next()
    return out


def check(q_tiles, n_tiles, R, cl, n_units, pinned):
    n_groups = (q_tiles + cl - 1) // cl
    n_items = ((n_tiles + R - 1) // R) * n_groups
    n_units = min(n_units, n_items)
    stride = n_groups * (n_units // n_groups) if (pinned and n_units >= n_groups) else 0
    seen = {}
    for u in range(n_units):
        vs = visits_of_unit(q_tiles, n_tiles, R, cl, u, n_units, stride)
        last_t = {}
        for t, g in vs:
            assert (t, g) not in seen, (t, g, u, seen.get((t, g)))
            seen[(t, g)] = u
            assert last_t.get(g, -1) < t                            # rows of a pool arrive in ascending order
            last_t[g] = t
        if stride:
            assert len({g for _, g in vs}) <= 1                     # pinned: one query-tile group per unit
    assert len(seen) == n_tiles * n_groups


def test_every_tile_group_pair_is_visited_exactly_once():
    for q_tiles, n_tiles, R, cl, n_units, pinned in itertools.product(
            (1, 2, 3, 9, 32), (1, 2, 7, 8, 33, 586), (1, 2, 3, 8), (1, 2, 4), (1, 5, 37, 74, 148), (False, True)):
        if R > n_tiles:
            continue
        check(q_tiles, n_tiles, R, cl, n_units, pinned)


def test_random_shapes():
    rng = random.Random(7)
    for _ in range(300):
        n_tiles = rng.randint(1, 3000)
        check(rng.randint(1, 32), n_tiles, rng.randint(1, min(8, n_tiles)), rng.choice((1, 2, 4, 8)),
              rng.randint(1, 148), rng.random() < 0.5)
