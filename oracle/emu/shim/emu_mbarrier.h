// emu_mbarrier.h - the four mbarrier / bulk-copy helpers of jmb_step_static.cuh (inline PTX on sm_100a) for the emulated
// library; make_emu.py puts this include where the PTX helpers stand, inside namespace jmb.  TEST INFRASTRUCTURE ONLY.
//   mbar_init(bar, count)            mbarrier.init: `count` arrivals per phase
//   mbar_expect_tx(bar, bytes)       mbarrier.arrive.expect_tx: one arrival + `bytes` of pending transactions
//   bulk_g2s(dst, src, bytes, bar)   cp.async.bulk ... mbarrier::complete_tx::bytes: the copy, then complete_tx(bytes)
//   mbar_wait(bar, parity)           mbarrier.try_wait.parity loop: returns once the phase of that parity has completed
// The copy is performed by the issuing thread and published through the barrier's lock, which is the visibility the
// hardware gives (data of a completed phase is visible to the waiters); re-use of a stage before every reader is done
// is then a data race ThreadSanitizer can see.
inline void mbar_init(uint64_t* bar, int count) { emu::mbar_init(bar, count); }
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { emu::mbar_arrive_expect_tx(bar, bytes); }
inline void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    std::memcpy(dst, src, bytes);
    emu::mbar_complete_tx(bar, bytes);
}
inline void mbar_wait(uint64_t* bar, uint32_t parity) { emu::mbar_wait(bar, parity); }
