""" Minimal stand-in for the third-party `batchflow` package (test infrastructure only).

Exists so that the UNMODIFIED reference module `/root/reference/pydens/model_torch.py`, which
hard-imports `batchflow.models.torch.Block` (model_torch.py:12) and star-imports
`batchflow.sampler` (pydens/__init__.py:5), can be imported in the build container.
All arithmetic stays PyTorch's / numpy's. """
