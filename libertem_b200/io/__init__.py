from .memory import MemoryDataSet, SyntheticDataSet, Partition

__all__ = ['MemoryDataSet', 'SyntheticDataSet', 'Partition']
