"""Step-size / momentum schedules: infinite iterators, as in optiml/opti/unconstrained/stochastic/schedules.py.
The device loop draws ``epochs`` values in advance and ships them as one array."""
import itertools


def constant(start):
    return itertools.repeat(start)


def decaying(start, decay):
    """start, start*decay, start*decay**2, ..."""
    return (start * decay ** i for i in itertools.count(0))


def linear_annealing(start, stop, n_steps):
    """n_steps values from start towards stop in equal increments, then stop for ever."""
    start, stop = float(start), float(stop)
    inc = (stop - start) / n_steps
    return itertools.chain((start + i * inc for i in range(n_steps)), itertools.repeat(stop))


def repeater(iterable, n):
    """every element of ``iterable`` n times in a row"""
    return (i for i in iterable for _ in range(n))
