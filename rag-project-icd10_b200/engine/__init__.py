"""Host side above the C ABI: vector table handle, record store, encoder engine, sharding."""
