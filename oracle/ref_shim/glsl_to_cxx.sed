# The one edit made to the reference's shader text before it is compiled as C++
# (oracle/Makefile, target `ref`): GLSL's array constructor
#     const mat4 m[4] = mat4[4] ( ... );      vec4 masks[4] = vec4[4] ( ... );
# becomes a C++ brace initialiser  = { ... };  Nothing else is touched.
s/= *mat4\[4\] *(/= {/
s/= *vec4\[4\] *(/= {/
s/^);[[:space:]]*$/};/
