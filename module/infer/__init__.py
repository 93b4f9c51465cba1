from tinyvc_b200.infer import BatchedStreamInfer, Generator, StreamInfer  # noqa: F401
