"""Mirror of BIOINF_tesi/visual for the one class that drives the predict path."""
from .visual import Compare_Models_Result, TASKS, CELL_LINES

__all__ = ['Compare_Models_Result', 'TASKS', 'CELL_LINES']
