"""Drop-in mirror of the reference's `lavila.models` modules for the dual-encoder hot path
(same class names, constructor arguments, parameter names/shapes and forward signatures)."""
