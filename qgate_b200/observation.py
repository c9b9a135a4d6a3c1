"""Observation containers returned by `Simulator.obs`, `Simulator.sample` and
`SamplingPool.sample` (public surface of qgate/simulator/observation.py:8-128)."""
import collections
import numbers


def _bits_repr(n_bits, value, none_mask):
    fmt = '{{:0{}b}}'.format(max(1, n_bits))
    return ''.join(v if m == '0' else '*'
                   for v, m in zip(fmt.format(int(value)), fmt.format(int(none_mask))))


class Observation:
    """One packed classical outcome; bit i belongs to reflist[i]."""

    def __init__(self, reflist, value, none_mask):
        self._reflist = reflist
        self._value = value
        self._none_mask = none_mask

    def __int__(self):
        return int(self._value)

    @property
    def int(self):
        return self._value

    def __call__(self, key):
        if key not in self._reflist:
            raise RuntimeError('unknown ref, {}.'.format(key))
        bit = 1 << self._reflist.index(key)
        if self._value & bit:
            return 1
        return None if self._none_mask & bit else 0

    def __eq__(self, other):
        if isinstance(other, numbers.Integral):
            raise ValueError('To compare with integer, use int() or Observation.int property.')
        if not isinstance(other, Observation):
            raise TypeError('comparison with Observation and {} not supported.'
                            .format(type(other)))
        return (self._reflist == other._reflist and self._value == other._value
                and self._none_mask == other._none_mask)

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = None

    def __repr__(self):
        return _bits_repr(len(self._reflist), self._value, self._none_mask)


class ObservationList:
    def __init__(self, reflist, values, mask):
        self._reflist = reflist
        self._values = values
        self._mask = mask

    @property
    def intarray(self):
        return self._values

    def __len__(self):
        return len(self._values)

    def __getitem__(self, key):
        if isinstance(key, slice):
            return ObservationList(self._reflist, self._values[key], self._mask)
        return Observation(self._reflist, self._values[key], self._mask)

    def __iter__(self):
        for value in self._values:
            yield Observation(self._reflist, value, self._mask)

    def __call__(self, ref):
        if ref not in self._reflist:
            raise RuntimeError('unknown ref, {}.'.format(ref))
        bit = 1 << self._reflist.index(ref)
        if self._mask & bit:
            return [None] * len(self._values)
        return [1 if (int(v) & bit) else 0 for v in self._values]

    def histgram(self):
        hist = collections.Counter(int(v) for v in self._values)
        return ObservationHistgram(hist, len(self._reflist), len(self))

    def __repr__(self):
        n_bits = len(self._reflist)
        shown = [_bits_repr(n_bits, v, self._mask) for v in self._values[:256]]
        if len(self) > 256:
            shown.append('...')
        return '[' + ', '.join(shown) + ']'


class ObservationHistgram:
    def __init__(self, hist, n_bits, n_samples):
        self._hist = hist
        self._n_bits = n_bits
        self._n_samples = n_samples

    @property
    def n_samples(self):
        return self._n_samples

    def __len__(self):
        return len(self._hist)

    def __iter__(self):
        return iter(self._hist)

    def __getitem__(self, key):
        return self._hist.get(key, 0)

    def keys(self):
        return self._hist.keys()

    def values(self):
        return self._hist.values()

    def items(self):
        return self._hist.items()

    def __repr__(self):
        fmt = '{{:0{}b}}: {{}}'.format(max(1, self._n_bits))
        return '{' + ', '.join(fmt.format(k, v) for k, v in sorted(self._hist.items())) + '}'
