from .bn import ABN, InPlaceABN, InPlaceABNSync, InPlaceABNWrapper, InPlaceABNSyncWrapper  # noqa: F401
