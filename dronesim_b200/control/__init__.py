"""Controller facades over the CUDA core (drop-in for ``dronesim.control``)."""
