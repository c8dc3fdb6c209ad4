"""Import shim so the reference (which needs only three timm symbols) can be imported offline."""
