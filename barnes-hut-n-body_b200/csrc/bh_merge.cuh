// bh_merge.cuh — the merge ("devour") rule of PhysicsEngine.step(), BarnesHutAlg.kt:463-532.
// Included by bh_engine.cu after the engine definition.
inline int bh_engine::merge_rule() {
    if (!merge_enabled()) return BH_OK;
    return fail(BH_E_UNSUPPORTED, "merge rule not implemented on device yet: set merge_min_dist <= 0");
}
