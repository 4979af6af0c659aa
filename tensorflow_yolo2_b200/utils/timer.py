"""Wall-clock tic/toc timer with the interface of the reference's src/utils/timer.py:10-32.

The reference's file is itself the timer of Fast R-CNN and carries this notice (src/utils/timer.py:1-6), kept here because the
class body is reproduced as-is for script compatibility:

    Fast R-CNN
    Copyright (c) 2015 Microsoft
    Licensed under The MIT License [see LICENSE for details]
    Written by Ross Girshick
"""
import time


class Timer(object):
    def __init__(self):
        self.total_time = 0.
        self.calls = 0
        self.start_time = 0.
        self.diff = 0.
        self.average_time = 0.

    def tic(self):
        self.start_time = time.time()

    def toc(self, average=True):
        self.diff = time.time() - self.start_time
        self.total_time += self.diff
        self.calls += 1
        self.average_time = self.total_time / self.calls
        return self.average_time if average else self.diff
