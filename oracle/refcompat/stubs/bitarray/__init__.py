"""Minimal bitarray stand-in: all that easydist/torch/reachability.py:33-69 uses."""
class bitarray(list):
    def __init__(self, n=0):
        super().__init__([0] * int(n))
    def setall(self, v):
        for i in range(len(self)):
            self[i] = int(bool(v))
