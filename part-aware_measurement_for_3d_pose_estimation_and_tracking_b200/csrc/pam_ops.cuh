// stateless batched ops (filled in below)
