"""Host-side mirror of the reference's ``air`` package (air/transformer.py, air/concrete.py,
air/vae.py, air/air_model.py): same function / constructor signatures, CUDA underneath."""
