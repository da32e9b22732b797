"""Drop-in mirror of the reference's `BIOINF_tesi` package for the EmbraceNet hot path (models + fit/predict).
Put this directory's parent on sys.path (or import `embrace_b200.BIOINF_tesi`) to switch an existing notebook to the
B200 engine.  `BIOINF_tesi.models`, the wire format of `data_pipe` and `visual.Compare_Models_Result` are provided; the rest of
data_pipe / visual stays with the reference."""
