"""Host adapters mirroring the reference's ``demo`` package (inference un-batching, evaluation summaries)."""
