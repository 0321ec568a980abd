# Stand-in for GD::Graph::histogram (only needed by bam2cfg.pl -h, which the goldens do not use).
package GD::Graph::histogram;
1;
