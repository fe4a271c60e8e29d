"""Stand-in for the `cached_property` PyPI package (reference requirements.txt:2),
which is not installable offline.  Test tooling only: lets the unmodified
reference under /root/reference import in this container."""
import functools

cached_property = functools.cached_property
