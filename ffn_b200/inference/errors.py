"""Exceptions of the inference stack (mirrors ffn/inference/errors.py)."""


class TerminationException(Exception):
  """Raised in client threads when the executor is shutting down."""
