from . import modeling  # noqa: F401  (registers heads / pixel decoders / transformer decoders)
