from .experience import Experience

__all__ = ["Experience"]
