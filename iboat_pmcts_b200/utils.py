"""Drop-in mirror of ``code/solver/utils.py`` (reference): the reward histogram ``Hist``.
(``Player``, the matplotlib widget of the reference, is out of scope.)

A ``Hist`` may own its 12 counters or be a *view* of one row of a tree's table
``table[node][8 actions][12 bins]`` -- the array-based trees of ``worker.Tree`` and the device-side
back-up kernel (``pmcts_backup``) work on such tables, and ``node.Values[i].h`` stays a live NumPy view.
"""
import numpy as np


def out_of_scope(name):
    """A callable named like one of the reference's plotting / download entry points: calling it says that the
    feature is outside the scope of this package instead of failing with an AttributeError."""
    def stub(*args, **kwargs):
        raise NotImplementedError("%s: the reference's plotting / animation / download code is outside the scope of "
                                  "pmcts_b200" % name)
    stub.__name__ = stub.__qualname__ = name
    return stub


Player = out_of_scope("utils.Player")


class Hist:
    #: Lower / upper bound of the histogram and number of bins (utils.py:14-20)
    MIN_REWARD = 0
    MAX_REWARD = 1.5
    N_BINS = 12
    #: Thresholds between bins, both ends included, and the centre of each bin (utils.py:23-26)
    THRESH = np.arange(MIN_REWARD, MAX_REWARD + (MAX_REWARD - MIN_REWARD) / N_BINS, (MAX_REWARD - MIN_REWARD) / N_BINS)
    MEANS = np.mean(np.stack((THRESH[:-1], THRESH[1:]), axis=0), axis=0)

    def __init__(self, init=[], store=None):
        if store is not None:
            self.h = store
        elif len(init) == 0:
            self.h = np.zeros(Hist.N_BINS, dtype=int)
        else:
            self.h = np.array(init, dtype=int)

    @staticmethod
    def bin_of(value):
        """Index of the bin ``add`` increments: the number of interior thresholds strictly below ``value``
        (values at or below the first threshold go to bin 0, above the last one to bin 11)."""
        i = 0
        for x in Hist.THRESH[1:-1]:
            if value > x:
                i = i + 1
            else:
                break
        return i

    def add(self, value):
        self.h[Hist.bin_of(value)] += 1

    def get_mean(self):
        summed = sum(self.h)
        if summed == 0:
            return 0
        return np.dot(self.h, Hist.MEANS) / summed

    def is_empty(self):
        return all(value == 0 for value in self.h)
