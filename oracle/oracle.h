// TEST INFRASTRUCTURE -- exported surface of the CPU oracle (liboracle.so). See oracle/README.md.
#pragma once
#include <stdint.h>
#include "../include/illuminant_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
int orc_render_lighting(const uint16_t* df_tex, int tw, int th, const void* gbuffer, int gw, int gh, int gfmt,
                        const ilb_lighting_frame* f, const ilb_light_batch* batches, int nbatches,
                        const ilb_light_vertex* verts, int nverts, float* out, int nthreads);
int orc_render_lighting_strided(const uint16_t* df_tex, int tw, int th, const void* gbuffer, int gw, int gh, int gfmt,
                                const ilb_lighting_frame* f, const ilb_light_batch* batches, int nbatches,
                                const ilb_light_vertex* verts, int nverts, int row_stride, float* out, int nthreads);
int orc_update_light_probes(const uint16_t* df_tex, int tw, int th, const ilb_lighting_frame* f,
                            const ilb_light_batch* batches, int nbatches, const ilb_light_vertex* verts, int nverts,
                            const ilb_float4* positions, const ilb_float4* normals, int nprobes, float* out);
float orc_sample_distance_field(const uint16_t* df_tex, int tw, int th, const ilb_df_uniforms* u, float x, float y, float z);
float orc_cone_trace(const uint16_t* df_tex, int tw, int th, const ilb_df_uniforms* u, const float* lightCenter,
                     float radius, float rampLength, float growth, float distanceFalloff, const float* shadedPos,
                     int enable, int* steps);
float orc_sphere_light_opacity(const ilb_lighting_frame* f, const float* pos, const float* normal, const float* center,
                               const float* lightProperties, float yFactor);
void orc_decode_gbuffer(const ilb_lighting_frame* f, const void* gbuffer, int gw, int gh, int gfmt, int x, int y,
                        float* worldPos, float* normal, int* enableShadows, int* fullbright);
float orc_evaluate_by_type_id(int typeId, const float* worldPosition, const float* center, const float* size, const float* rotation);
/* LifeRampTexture of the particle update (UpdateCommon.fxh:6-13): float4 texels, row-major; NULL = none. The pointer must stay valid. */
void orc_set_life_ramp(const float* texels, int w, int h);
float orc_bezier1(const ilb_bezier1* b, float value);
void orc_bezier4(const ilb_bezier4* b, float value, float* out);
int orc_particles_step(float* P, float* V, float* A, float* RC, float* RD, int chunk_size, int live_chunks,
                       const ilb_psys_uniforms* u, const ilb_spawn* spawns, int nspawns, const ilb_op* ops, int nops,
                       const float* rng_table, int rw, int rh, const uint16_t* df_tex, int tw, int th, int steps,
                       int nthreads);
/* N4: spawners that read a source -- `sources` as for ilb_particles_step_sources (source_system is ignored); for FEEDBACK entries
 * states[i] holds the SOURCE CHUNK's PositionAndLife / Velocity / RenderColor (chunk_size^2 float4 each). */
typedef struct orc_source_state { const float *P, *V, *RC; int chunk_size; } orc_source_state;
int orc_particles_step_sources(float* P, float* V, float* A, float* RC, float* RD, int chunk_size, int live_chunks,
                               const ilb_psys_uniforms* u, const ilb_spawn* spawns, const ilb_spawn_source* sources,
                               const orc_source_state* states, int nspawns, const ilb_op* ops, int nops,
                               const float* rng_table, int rw, int rh, const uint16_t* df_tex, int tw, int th, int steps,
                               int nthreads);
/* N2: ParticleSystem.Render over `total` particles in draw order; target: width*height float4 (read when clear == 0, written). */
int orc_particles_render(const float* P, const float* RD, const float* RC, long total, const ilb_particle_render* r,
                         const uint8_t* texture, float* target);
int orc_generate_distance_field(uint16_t* out_rgba64, const uint16_t* base_rgba64 /* static field or NULL */, int tw, int th, int slice_w, int slice_h, int slice_count,
                                const ilb_df_uniforms* u, const ilb_obstruction* obs, int count, int nthreads);
/* The same for physical slices [first_physical, first_physical + physical_count) only, IN PLACE on `tex` (other slices keep their
 * texels), with height volumes (Shaders/DistanceField.fx, see ilb_height_volume). */
int orc_update_distance_field_slices(uint16_t* tex, const uint16_t* base, int tw, int th, int slice_w, int slice_h, int slice_count,
                                     const ilb_df_uniforms* u, const ilb_obstruction* obs, int count, const ilb_height_volume* volumes,
                                     int nvolumes, const ilb_float4* edges, int nedges, int first_physical, int physical_count, int nthreads);
/* signed distance of one point to a height volume at slice depth z (computeSliceDistances / finalEval of DistanceField.fx) */
float orc_height_volume_distance(const ilb_height_volume* volume, const ilb_float4* edges, float x, float y, float z);
void orc_encode_gbuffer_sample(const float* normal, float relativeY, float z, int dead, int enableShadows, int fullbright, float* out4);
/* N3: Resolve.fx / HDR.fxh on fp32-decoded texels (lightmap, albedo: w*h*4 floats; albedo may be NULL); out: w*h*4 floats, not quantised. */
int orc_resolve_lighting(const ilb_resolve* p, const float* lightmap, const float* albedo, float* out);
/* N3: the resolve drawn as a quad at placement->Position with placement->Scale into a target of any size (ResolveLighting,
 * LightingRenderer.cs:1537-1645): lightmap w*h*4 floats, albedo (nullable) albedo_width*albedo_height*4 floats, target
 * target_width*target_height*4 floats read and written (pixels outside the quad keep their contents). */
/* SphereLightWithDistanceRamp: the ramp textures (float4 texels, kept by reference) that ilb_light_batch.ramp_texture = id
 * (1-based) selects in later orc_render_lighting calls. */
void orc_set_ramp_textures(int count, const float* const* texels, const int* widths, const int* heights);
/* N3: ApplyDither settings of every later orc_resolve_* call (NULL = the handler's default, Strength 0); LUT-blended resolve
 * (LUTResolve.fx); textures are float4 texels. */
void orc_set_dithering(const ilb_dithering* d);
int orc_resolve_lighting_lut(const ilb_resolve* p, const ilb_lut_blending* lut, const float* dark, const float* bright, const float* lightmap,
                             const float* albedo, float* out);
int orc_resolve_lighting_placed(const ilb_resolve* p, const ilb_resolve_placement* place, const float* lightmap, const float* albedo, float* target);
/* N3: luminance buffer level `level` ((w/2 >> level) x (h/2 >> level) floats) of a fp32-decoded lightmap. */
int orc_compute_luminance(const float* lightmap, int w, int h, int level, float* out);
/* The host build of include/ilb_detmath.h (what every oracle function above calls): function 0 dm_sinf, 1 dm_cosf, 2 dm_acosf. */
void orc_detmath(int function, const float* x, float* out, long n);
void orc_float_to_half(const float* in, uint16_t* out, long n);
void orc_half_to_float(const uint16_t* in, float* out, long n);
#ifdef __cplusplus
}
#endif
