"""Drop-in mirror of the reference's `BIOINF_tesi` package for the EmbraceNet hot path (models + fit/predict).
Put this directory's parent on sys.path (or import `embrace_b200.BIOINF_tesi`) to switch an existing notebook to the
B200 engine.  Only `BIOINF_tesi.models` is provided: data_pipe / visual stay with the reference."""
