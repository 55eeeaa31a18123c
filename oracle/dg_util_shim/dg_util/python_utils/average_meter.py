"""Import-level stand-in for dg_util.python_utils.average_meter (TEST INFRASTRUCTURE ONLY): the two meters
solvers/vince_solver.py and solvers/base_solver.py update every iteration."""


class AverageMeter(object):
    def __init__(self):
        self.reset()

    def reset(self):
        self.val, self.avg, self.sum, self.count = 0, 0, 0, 0

    def update(self, val, n=1):
        val = float(val)
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


class RollingAverageMeter(object):
    def __init__(self, window_size=10):
        self.window_size = window_size
        self.reset()

    def reset(self):
        self.history = []
        self.val = 0

    def update(self, val, n=1):
        self.history.append(float(val))
        self.history = self.history[-self.window_size:]
        self.val = sum(self.history) / len(self.history)
