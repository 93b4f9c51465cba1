"""Import-path shim: `module.*` resolves to the CUDA implementation in `tinyvc_b200` so that the
reference's entry points and user scripts (`from module.tinyvc import Encoder`) are drop-in."""
