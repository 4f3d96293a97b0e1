"""Drop-in mirrors of the reference's ``scripts`` modules (same names and signatures)."""
