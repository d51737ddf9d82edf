"""Result containers of the analysis wrappers (raw data only; rendering -- the reference's
``visualized`` channel, common/analysis.py -- is out of scope of the hot path)."""


class AnalysisResult:
    def __init__(self, raw_data, key, title=None, desc=None):
        self.raw_data = raw_data
        self.key = key
        self.title = title or key
        self.desc = desc or key

    def __array__(self, dtype=None, copy=None):
        import numpy as np
        return np.asarray(self.raw_data, dtype=dtype)

    def __repr__(self):
        return f'<AnalysisResult: {self.key}>'


class AnalysisResultSet:
    def __init__(self, results, raw_results=None):
        self._results = results
        self.raw_results = raw_results

    @property
    def results(self):
        if callable(self._results):
            self._results = self._results()
        return self._results

    def __getattr__(self, k):
        if k.startswith('_'):
            raise AttributeError(k)
        for r in self.results:
            if r.key == k:
                return r
        raise AttributeError(k)

    def __getitem__(self, k):
        if isinstance(k, str):
            return getattr(self, k)
        return self.results[k]

    def keys(self):
        return [r.key for r in self.results]

    def __len__(self):
        return len(self.results)

    def __iter__(self):
        return iter(self.results)
