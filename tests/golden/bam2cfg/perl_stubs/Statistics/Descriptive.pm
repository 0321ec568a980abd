# Minimal stand-in for the CPAN module Statistics::Descriptive (absent from this image): only what perl/bam2cfg.pl calls
# (add_data, get_data, count, mean, standard_deviation), with the module's formulas (running sum / sum of squares, n - 1).
# Test infrastructure: lets the UNMODIFIED reference script run here to produce tests/golden/bam2cfg/*.cfg.
package Statistics::Descriptive;
1;
package Statistics::Descriptive::Full;
sub new { my $c = shift; return bless { data => [], sum => 0, sumsq => 0 }, $c; }
sub add_data { my $s = shift; for my $v (@_) { push @{$s->{data}}, $v; $s->{sum} += $v; $s->{sumsq} += $v * $v; } return 1; }
sub get_data { my $s = shift; return @{$s->{data}}; }
sub count { my $s = shift; return scalar @{$s->{data}}; }
sub mean { my $s = shift; my $n = $s->count(); return undef unless $n; return $s->{sum} / $n; }
sub variance { my $s = shift; my $n = $s->count(); return undef unless $n; return 0 if $n < 2;
               my $m = $s->mean(); my $v = ($s->{sumsq} - $n * $m * $m) / ($n - 1); return $v < 0 ? 0 : $v; }
sub standard_deviation { my $s = shift; my $v = $s->variance(); return undef unless defined $v; return sqrt($v); }
1;
